mkdir -p gpurun_out/r02f
(timeout 1700 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02f/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02f/pytest.log); tail -15 gpurun_out/r02f/pytest.log | cut -c1-250
for sp in "" 2; do
B200_TILE_SPLIT=$sp BENCH_NO_CLOCKS=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-configs --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('split=$sp', d['ms_per_step'], d['stage_ms'])" | tee -a gpurun_out/r02f/ab.txt
done
