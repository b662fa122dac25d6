mkdir -p gpurun_out/r02g
(timeout 1200 python -m pytest tests/test_gpu_lammps_shell.py tests/test_gpu_parity.py -m gpu -q --timeout 900 -k "shell or lj or 1000 or refuses" > gpurun_out/r02g/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02g/pytest.log); tail -8 gpurun_out/r02g/pytest.log | cut -c1-300
BENCH_NO_CLOCKS=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-configs --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])" | tee -a gpurun_out/r02g/ab.txt
