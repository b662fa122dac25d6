# 1 GPU: fix phase_change with the skip-ahead RNG walk: every phase-change fixture (exact insertion sequences), the 1 M oracle check, C4 timings
mkdir -p gpurun_out/r02v
(timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -k "bubble or kat_phase or phase_change or c4" > gpurun_out/r02v/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02v/pytest.log); tail -5 gpurun_out/r02v/pytest.log | cut -c1-300
echo "== c4 1M"; timeout 300 python tests/dev_bench.py c4 100 20 2>&1 | grep -E "ms/step|phase_change|neigh_bin" | cut -c1-160
echo "== c4 4M"; timeout 300 python tests/dev_bench.py c4 160 10 2>&1 | grep -E "ms/step|phase_change|neigh_bin" | cut -c1-160
