# 1 GPU A/B: near writer with class counters in shared memory (n1) vs in registers (n0); density pass with 2 lanes per row and two CTAs per SM vs 4 lanes
mkdir -p gpurun_out/r02p
AB=$PWD/lammps-sph-multiphase_b200/csrc/ab
for v in n1 n0 n1; do
B200_LIB=$AB/$v.so BENCH_NO_CLOCKS=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-configs --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['ms_per_step'], d['stage_ms'])"
done
for ds in 2 4; do
B200_DENSITY_SPLIT=$ds BENCH_NO_CLOCKS=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-configs --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dsplit=$ds', d['ms_per_step'], d['stage_ms'])"
done
for v in n1 n0; do echo "== c3 $v"; B200_LIB=$AB/$v.so timeout 300 python tests/dev_bench.py c3 100 20 2>&1 | grep -E "ms/step|neigh_build" | cut -c1-120; done
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tile.py tests/test_gpu_edge.py -m gpu -q --timeout 600 > gpurun_out/r02p/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02p/pytest.log); tail -4 gpurun_out/r02p/pytest.log | cut -c1-300
