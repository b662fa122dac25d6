# 1 GPU: full bench line (C2 + C3 + C5 at N=1), then the launch list and one full capture of the stage kernels
mkdir -p gpurun_out/r02d
timeout 900 python bench.py > gpurun_out/r02d/bench.json 2> gpurun_out/r02d/bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02d/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02d/bench.json'))
print(d['ms_per_step'], d['stage_ms'], d['e2e']['value'], d['roofline']['frac'])
for k,v in d['configs'].items(): print(k, {a:v.get(a) for a in ('ms_per_step','particle_steps_s','particles_total','stage_ms','sum_f_over_sum_abs_f','setup_seconds','error')})
PY
S="python bench.py --steps 10 --warmup 3 --no-configs --no-e2e --no-cpu-baseline"
timeout 300 $S > gpurun_out/r02d/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02d/launches.csv $S > gpurun_out/r02d/ncu1.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_tile_(force|rhosum)' -s 4 -c 2 -o gpurun_out/r02d/stage -f $S > gpurun_out/r02d/ncu2.log 2>&1; echo "ncu full rc=$?"
