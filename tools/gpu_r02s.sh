# final-state validation on 1 GPU: smoke(), the whole GPU suite, the driver-style bench line and a 200-step line, the reference arm
mkdir -p gpurun_out/r02s
python __graft_entry__.py smoke 2>&1 | tail -3
(timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02s/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02s/pytest.log); tail -6 gpurun_out/r02s/pytest.log | cut -c1-300
(time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02s/bench_driver.json 2> gpurun_out/r02s/bench.err) 2>&1 | grep real; tail -c 300 gpurun_out/r02s/bench.err
timeout 600 python bench.py --no-configs > gpurun_out/r02s/bench_200.json 2>> gpurun_out/r02s/bench.err
python - <<'PY'
import json
for f in ('bench_driver','bench_200'):
    d=json.load(open('gpurun_out/r02s/%s.json'%f))
    print(f, d['ms_per_step'], d['value'], d['stage_ms'], 'e2e', d['e2e']['value'], d['e2e'].get('breakdown_s'), 'roof', d['roofline']['frac'], d['roofline'].get('fp64_frac'), d['clocks'])
    for k,v in d.get('configs',{}).items(): print(' ', k, {a:v.get(a) for a in ('ms_per_step','particle_steps_s','particles_total','setup_seconds','error')})
PY
(time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02s/bench_ref.json 2>> gpurun_out/r02s/bench.err) 2>&1 | grep real; cut -c1-400 gpurun_out/r02s/bench_ref.json
