# half_bin_newton ownership (k_build), stale-vest setups on the row path, elapsed time under fix dt/reset: the fixture parity tests
# without the 1000-step decks (the GPU budget of the round was down to two minutes)
mkdir -p gpurun_out/hbn
(timeout 80 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 60 -p no:cacheprovider -k "not 1000" > gpurun_out/hbn/parity.log 2>&1; echo "rc=$?" >> gpurun_out/hbn/parity.log)
tail -30 gpurun_out/hbn/parity.log | cut -c1-300
