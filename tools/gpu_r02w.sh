# 1 GPU: three-launch scan: parity + timings of the ghost construction stage
mkdir -p gpurun_out/r02w
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tile.py tests/test_gpu_edge.py -m gpu -q --timeout 600 > gpurun_out/r02w/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02w/pytest.log); tail -4 gpurun_out/r02w/pytest.log | cut -c1-300
echo "== c3 1M"; timeout 300 python tests/dev_bench.py c3 100 20 2>&1 | grep -E "ms/step|neigh_bin|neigh_build" | cut -c1-160
echo "== c4 1M"; timeout 300 python tests/dev_bench.py c4 100 20 2>&1 | grep -E "ms/step|neigh_bin|phase_change" | cut -c1-160
BENCH_NO_CLOCKS=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-configs --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])"
