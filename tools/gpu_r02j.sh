# 1 GPU A/B of compile-time variants (lammps-sph-multiphase_b200/csrc/ab/*.so, selected with B200_LIB): row writers, flattened emission, L2 prefetch
mkdir -p gpurun_out/r02j
for v in v1 v0 v2 p0 v1; do
B200_LIB=$PWD/lammps-sph-multiphase_b200/csrc/ab/$v.so BENCH_NO_CLOCKS=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-configs --no-e2e --no-cpu-baseline 2>gpurun_out/r02j/bench_$v.err > gpurun_out/r02j/bench_$v.json
python -c "import json,sys; d=json.loads(open('gpurun_out/r02j/bench_$v.json').read()); print('$v', d['ms_per_step'], d['stage_ms'])"
done
for v in v1 p0 v0; do
echo "== c3 $v"; B200_LIB=$PWD/lammps-sph-multiphase_b200/csrc/ab/$v.so timeout 300 python tests/dev_bench.py c3 100 20 2>&1 | tail -12 | cut -c1-200
done
echo "== c4 v1"; B200_LIB=$PWD/lammps-sph-multiphase_b200/csrc/ab/v1.so timeout 300 python tests/dev_bench.py c4 100 20 2>&1 | tail -12 | cut -c1-200
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tile.py -m gpu -q -x --timeout 600 > gpurun_out/r02j/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02j/pytest.log); tail -4 gpurun_out/r02j/pytest.log | cut -c1-300
