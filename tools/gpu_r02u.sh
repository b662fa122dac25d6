# 2 GPUs: decomposition parity + the driver-style bench line with complete-tank weak-scaling tiles
mkdir -p gpurun_out/r02u
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29511 tests/mgpu_check.py > gpurun_out/r02u/mgpu_check.log 2>&1; echo "mgpu rc=$?"; grep -c " OK " gpurun_out/r02u/mgpu_check.log; grep -E "FAIL|Error" gpurun_out/r02u/mgpu_check.log | head -5
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 500 2>&1 | tail -3
timeout 900 $T --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02u/bench_n2.json 2> gpurun_out/r02u/bench_n2.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02u/bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02u/bench_n2.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['value'], d['config']['particles_total'], d['stage_ms'], 'e2e', d['e2e']['value'], d['e2e'].get('breakdown_s'))
for k,v in d['configs'].items(): print(k, {a:v.get(a) for a in ('ms_per_step','particle_steps_s','particles_total','grid','error')})
print(d['parity']['ok'], d['parity']['max_err'])
PY
