# 1 GPU: fused border flags/scan: parity (periodic decks exercise the self swaps) + timings; full captures of the multiphase density / colorgradient kernels
mkdir -p gpurun_out/r02x
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tile.py tests/test_gpu_edge.py -m gpu -q --timeout 600 > gpurun_out/r02x/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02x/pytest.log); tail -3 gpurun_out/r02x/pytest.log | cut -c1-300
echo "== c3 1M"; timeout 300 python tests/dev_bench.py c3 100 20 2>&1 | grep -E "ms/step|neigh_bin|neigh_build" | cut -c1-160
S="python tests/dev_bench.py c3 100 2"
ncu --set full --clock-control none --import-source on -k regex:'k_tile_full_mp' -s 4 -c 2 -o gpurun_out/r02x/mpfull -f $S > gpurun_out/r02x/ncu1.log 2>&1; echo "ncu rc=$?"
