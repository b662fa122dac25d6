# 1 GPU: build without zone compares on skin-0 decks: parity + C3 / C4 timings
mkdir -p gpurun_out/r02y
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tile.py tests/test_gpu_edge.py -m gpu -q --timeout 600 > gpurun_out/r02y/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02y/pytest.log); tail -3 gpurun_out/r02y/pytest.log | cut -c1-300
echo "== c3 1M"; timeout 300 python tests/dev_bench.py c3 100 20 2>&1 | grep -E "ms/step|neigh_bin|neigh_build" | cut -c1-160
echo "== c4 1M"; timeout 300 python tests/dev_bench.py c4 100 20 2>&1 | grep -E "ms/step|neigh_build" | cut -c1-160
