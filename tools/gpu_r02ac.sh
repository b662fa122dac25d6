# 2 GPUs: exchange with the reference's hole-filling order: the default decomposition checks, then the order-sensitive deck against the P-rank oracle
mkdir -p gpurun_out/r02ac
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29511 tests/mgpu_check.py > gpurun_out/r02ac/mgpu_check.log 2>&1; echo "mgpu rc=$?"; grep " grid " gpurun_out/r02ac/mgpu_check.log | cut -c1-150
timeout 300 $T --master-port 29512 tests/mgpu_check.py --vs-world shock3d shock2d dam3d dam2d > gpurun_out/r02ac/mgpu_world.log 2>&1; echo "world rc=$?"; grep " grid " gpurun_out/r02ac/mgpu_world.log | cut -c1-330
timeout 300 $T --master-port 29513 tests/mgpu_check.py --empty-rank > gpurun_out/r02ac/mgpu_empty.log 2>&1; echo "empty rc=$?"
