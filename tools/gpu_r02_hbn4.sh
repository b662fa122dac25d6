# last short validation of the round: pair_style hybrid with `pair_coeff 2 3 none` (cutneighsq 0 for a type pair, hybrid map) on the engine
mkdir -p gpurun_out/hbn
(timeout 25 python -m pytest -m gpu -q --timeout 20 -p no:cacheprovider "tests/test_gpu_parity.py::test_engine_matches_reference_fixture[cavity2d_none]" > gpurun_out/hbn/none.log 2>&1; echo "rc=$?" >> gpurun_out/hbn/none.log)
tail -30 gpurun_out/hbn/none.log | cut -c1-400
