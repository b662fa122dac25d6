# 1 GPU: parity after the multiphase force reformulation (surface-stress tensor per particle) and the pipelined exact tests of the build; C2 / C3 / C4 timings
mkdir -p gpurun_out/r02l
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tile.py tests/test_gpu_edge.py -m gpu -q --timeout 600 > gpurun_out/r02l/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02l/pytest.log); tail -12 gpurun_out/r02l/pytest.log | cut -c1-400
BENCH_NO_CLOCKS=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-configs --no-e2e --no-cpu-baseline 2>gpurun_out/r02l/bench.err > gpurun_out/r02l/bench.json
python -c "import json,sys; d=json.loads(open('gpurun_out/r02l/bench.json').read()); print(d['ms_per_step'], d['stage_ms'])"
echo "== c3"; timeout 300 python tests/dev_bench.py c3 100 20 2>&1 | tail -10 | cut -c1-200
echo "== c4"; timeout 300 python tests/dev_bench.py c4 100 20 2>&1 | tail -11 | cut -c1-200
