# final code, 1 GPU: smoke(), the whole GPU suite, the driver-style bench line (with C3 / C4 / C5) and the reference arm
mkdir -p gpurun_out/final1
python __graft_entry__.py smoke 2>&1 | tail -2
(timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/final1/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/final1/pytest.log); tail -4 gpurun_out/final1/pytest.log | cut -c1-300
(time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/final1/bench_driver.json 2> gpurun_out/final1/bench.err) 2>&1 | grep real; tail -c 300 gpurun_out/final1/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/final1/bench_driver.json'))
print(d['ms_per_step'], d['value'], d['stage_ms'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline'].get('fp64_frac'), d['gpu_launches'])
for k,v in d.get('configs',{}).items(): print(' ', k, {a:v.get(a) for a in ('ms_per_step','particle_steps_s','particles_total','inserted','error')})
PY
