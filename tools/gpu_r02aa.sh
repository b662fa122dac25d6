# 1 GPU A/B: unroll depth of the force bodies (single-phase 8 vs 4, multiphase 4 vs 2 vs 8); warp-per-tile k_tile_zone is in all of them
AB=$PWD/lammps-sph-multiphase_b200/csrc/ab
for v in base f4m2 base; do
B200_LIB=$AB/$v.so BENCH_NO_CLOCKS=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-configs --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['ms_per_step'], d['stage_ms'])"
done
for v in base f4m2 m8; do echo "== c3 $v"; B200_LIB=$AB/$v.so timeout 300 python tests/dev_bench.py c3 100 20 2>&1 | grep -E "ms/step|force " | cut -c1-120; done
(timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "zone or dam" 2>&1 | tail -2)
