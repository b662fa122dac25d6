# final code, 1 GPU: launch list of the bench command (ncu, serialised) + full capture of the stage kernels + bench lines
mkdir -p gpurun_out/r02z
S="python bench.py --steps 10 --warmup 3 --no-configs --no-e2e --no-cpu-baseline"
timeout 300 $S > gpurun_out/r02z/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z/launches.csv $S > gpurun_out/r02z/ncu1.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_tile_(force|rhosum|build)' -s 6 -c 3 -o gpurun_out/r02z/stage -f $S > gpurun_out/r02z/ncu2.log 2>&1; echo "ncu full rc=$?"
(time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02z/bench_driver.json 2> gpurun_out/r02z/bench.err) 2>&1 | grep real
timeout 600 python bench.py --no-configs > gpurun_out/r02z/bench_200.json 2>> gpurun_out/r02z/bench.err
python - <<'PY'
import json
for f in ('bench_driver','bench_200'):
    d=json.load(open('gpurun_out/r02z/%s.json'%f))
    print(f, d['ms_per_step'], d['value'], d['stage_ms'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline'].get('fp64_frac'))
    for k,v in d.get('configs',{}).items(): print(' ', k, {a:v.get(a) for a in ('ms_per_step','particle_steps_s','particles_total','error')})
PY
