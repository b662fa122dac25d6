# 1 GPU: density pass with two CTAs per SM as the default; parity
mkdir -p gpurun_out/r02q
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tile.py tests/test_gpu_edge.py -m gpu -q --timeout 600 > gpurun_out/r02q/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02q/pytest.log); tail -4 gpurun_out/r02q/pytest.log | cut -c1-300
for i in 1 2; do
BENCH_NO_CLOCKS=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-configs --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])"
done
