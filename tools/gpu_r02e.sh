# 2 GPUs: decomposition parity (fixtures / P-rank oracle / empty brick), then the bench line at N = 2
mkdir -p gpurun_out/r02e
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29511 tests/mgpu_check.py > gpurun_out/r02e/mgpu_check.log 2>&1; echo "mgpu rc=$?"; grep " grid " gpurun_out/r02e/mgpu_check.log | cut -c1-230
timeout 300 $T --master-port 29512 tests/mgpu_check.py --empty-rank > gpurun_out/r02e/mgpu_empty.log 2>&1; echo "empty rc=$?"; grep " grid " gpurun_out/r02e/mgpu_empty.log | cut -c1-230; tail -3 gpurun_out/r02e/mgpu_empty.log | cut -c1-300
B200_OVERLAP=1 timeout 300 $T --master-port 29513 tests/mgpu_check.py dam3d heat3d > gpurun_out/r02e/mgpu_overlap.log 2>&1; echo "overlap rc=$?"; grep " grid " gpurun_out/r02e/mgpu_overlap.log | cut -c1-200
timeout 900 $T --master-port 29514 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r02e/bench_n2.json 2> gpurun_out/r02e/bench_n2.err; echo "bench rc=$?"; tail -c 500 gpurun_out/r02e/bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02e/bench_n2.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['value'], d['stage_ms'], d['e2e']['value'])
for k,v in d['configs'].items(): print(k, {a:v.get(a) for a in ('ms_per_step','particle_steps_s','particles_total','grid','stage_ms','sum_f_over_sum_abs_f','inserted','error')})
print(d['parity'])
PY
