# 1 GPU: the whole GPU suite (new: fix addforce / setmeso with variables, merged ghost-row tiles), then the default bench line as the driver runs it
mkdir -p gpurun_out/r02n
(timeout 1700 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02n/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02n/pytest.log); tail -15 gpurun_out/r02n/pytest.log | cut -c1-400
(time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02n/bench.json 2> gpurun_out/r02n/bench.err) 2>&1 | grep real; tail -c 300 gpurun_out/r02n/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02n/bench.json'))
print(d['ms_per_step'], d['value'], d['stage_ms'], 'e2e', d['e2e']['value'], d['e2e'].get('breakdown_s'), 'roof', d['roofline']['frac'], d['roofline'].get('fp64_frac'))
for k,v in d['configs'].items(): print(k, {a:v.get(a) for a in ('ms_per_step','particle_steps_s','particles_total','stage_ms','sum_f_over_sum_abs_f','setup_seconds','error')})
print(d.get('cpu_baseline'))
PY
