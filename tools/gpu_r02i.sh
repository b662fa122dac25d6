# 1 GPU: parity after the build rewrite (flattened emission), A/B of the force kernel's lanes per row, multiphase dev bench
mkdir -p gpurun_out/r02i
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tile.py tests/test_gpu_edge.py -m gpu -q -x --timeout 600 > gpurun_out/r02i/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02i/pytest.log); tail -6 gpurun_out/r02i/pytest.log | cut -c1-300
for fs in 2 4; do
B200_FORCE_SPLIT=$fs BENCH_NO_CLOCKS=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-configs --no-e2e --no-cpu-baseline 2>gpurun_out/r02i/bench_fs$fs.err > gpurun_out/r02i/bench_fs$fs.json
python -c "import json,sys; d=json.loads(open('gpurun_out/r02i/bench_fs$fs.json').read()); print('fsplit=$fs', d['ms_per_step'], d['stage_ms'])"
done
timeout 300 python tests/dev_bench.py c3 100 20 2>&1 | tail -4 | cut -c1-400
timeout 300 python tests/dev_bench.py c4 100 20 2>&1 | tail -4 | cut -c1-400
