# 1 GPU: parity of the new force body, bench at the driver's step count and at 200 steps (e2e breakdown), source-level capture of the build
mkdir -p gpurun_out/r02h
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tile.py -m gpu -q -x --timeout 600 > gpurun_out/r02h/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02h/pytest.log); tail -6 gpurun_out/r02h/pytest.log | cut -c1-300
for K in "20 5" "200 20"; do set -- $K
BENCH_NO_CLOCKS=1 timeout 300 python bench.py --steps $1 --warmup $2 --no-configs --no-cpu-baseline 2>gpurun_out/r02h/bench_$1.err > gpurun_out/r02h/bench_$1.json
python -c "import json,sys; d=json.loads(open('gpurun_out/r02h/bench_$1.json').read()); print('steps=$1', d['ms_per_step'], d['stage_ms'], 'e2e', d['e2e']['value'], d['e2e']['seconds'], d['e2e'].get('breakdown_s'))"
done
S="python bench.py --steps 6 --warmup 3 --no-configs --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:'k_tile_build' -s 1 -c 1 -o gpurun_out/r02h/build -f $S > gpurun_out/r02h/ncu1.log 2>&1; echo "ncu build rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_tile_force' -s 4 -c 1 -o gpurun_out/r02h/force -f $S > gpurun_out/r02h/ncu2.log 2>&1; echo "ncu force rc=$?"
ls -la gpurun_out/r02h
