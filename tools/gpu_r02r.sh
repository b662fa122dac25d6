# 1 GPU: run-to-run variance of the build stage (row-stride head room 1.4x), 4 runs of 100 steps + 1 of 300
for i in 1 2 3 4; do
B200_VERBOSE=1 BENCH_NO_CLOCKS=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-configs --no-e2e --no-cpu-baseline 2>gpurun_out/err.txt | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d['config']['l2'][:60])"; grep "b200:" gpurun_out/err.txt | head -5
done
B200_VERBOSE=1 BENCH_NO_CLOCKS=1 timeout 300 python bench.py --steps 300 --warmup 10 --no-configs --no-e2e --no-cpu-baseline 2>gpurun_out/err.txt | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d['config']['l2'][:60])"; grep "b200:" gpurun_out/err.txt | head -5
