# final code, 4 GPUs: the driver's launch of the bench
mkdir -p gpurun_out/final4
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
(time timeout 900 $T --master-port 29621 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/final4/bench_n4.json 2> gpurun_out/final4/bench_n4.err) 2>&1 | grep real; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/final4/bench_n4.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['value'], d['config']['particles_total'], d['stage_ms'], 'e2e', d['e2e']['value'])
for k,v in d['configs'].items(): print(k, {a:v.get(a) for a in ('ms_per_step','particle_steps_s','particles_total','grid','error')}, {a:round(b/v['steps'],2) for a,b in (v.get('stage_ms') or {}).items()} if v.get('stage_ms') else '')
print(d['parity']['ok'], d['parity']['max_err'], d['parity']['grid'])
PY
