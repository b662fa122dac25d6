# 8 GPUs: the driver's launch of the bench (C2 weak + C5 strong on the 2x2x2 grid + one-tank balance + parity on the same ranks)
mkdir -p gpurun_out/r02t
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
(time timeout 1200 $T --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02t/bench_n8.json 2> gpurun_out/r02t/bench_n8.err) 2>&1 | grep real; echo "bench rc=$?"; tail -c 600 gpurun_out/r02t/bench_n8.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02t/bench_n8.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['value'], d['stage_ms'], 'e2e', d['e2e']['value'], d['e2e'].get('breakdown_s'))
for k,v in d['configs'].items(): print(k, {a:v.get(a) for a in ('ms_per_step','particle_steps_s','particles_total','grid','stage_ms','sum_f_over_sum_abs_f','inserted','error','uniform','balanced')})
print(d['parity'])
PY
