# final code, 2 GPUs: local order / decomposition parity against the P-rank oracle, then the driver-style bench line
mkdir -p gpurun_out/final2
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $T --master-port 29512 tests/mgpu_check.py --vs-world dam3d dam2d_1000 droplet2d bubble2d_1000 shock3d > gpurun_out/final2/mgpu_world.log 2>&1; echo "world rc=$?"; grep " grid " gpurun_out/final2/mgpu_world.log | cut -c1-200
timeout 400 $T --master-port 29511 tests/mgpu_check.py > gpurun_out/final2/mgpu_check.log 2>&1; echo "mgpu rc=$?"; grep -c " OK " gpurun_out/final2/mgpu_check.log
timeout 900 $T --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/final2/bench_n2.json 2> gpurun_out/final2/bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/final2/bench_n2.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['value'], d['config']['particles_total'], d['stage_ms'], 'e2e', d['e2e']['value'])
for k,v in d['configs'].items(): print(k, {a:v.get(a) for a in ('ms_per_step','particle_steps_s','particles_total','grid','error')}, {a:round(b/v['steps'],2) for a,b in (v.get('stage_ms') or {}).items()} if v.get('stage_ms') else '')
print(d['parity']['ok'], d['parity']['max_err'])
PY
