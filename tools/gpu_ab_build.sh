mkdir -p gpurun_out/r02c
B="python bench.py --steps 50 --warmup 10 --no-configs --no-e2e --no-cpu-baseline"
for nb in 32 64; do for nt in 128 256; do
  B200_BUILD_NBLK=$nb B200_BUILD_NT=$nt BENCH_NO_CLOCKS=1 timeout 300 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C2 nblk=$nb nt=$nt', d['ms_per_step'], d['stage_ms'])" | tee -a gpurun_out/r02c/build_ab.txt
  B200_BUILD_NBLK=$nb B200_BUILD_NT=$nt timeout 300 python tests/dev_bench.py c3 100 20 2>&1 | grep -E "ms/step|neigh_build|force|density|colorgradient" | sed "s/^/C3 nblk=$nb nt=$nt /" | tee -a gpurun_out/r02c/build_ab.txt
done; done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tile.py tests/test_gpu_edge.py tests/test_gpu_fullsize.py -m gpu -q --timeout 600 > gpurun_out/r02c/pytest.log 2>&1; tail -3 gpurun_out/r02c/pytest.log
