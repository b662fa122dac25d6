# 2 GPUs: local order of every rank after migration against the P-rank oracle (dam break: many atoms cross the cut)
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29512 tests/mgpu_check.py --vs-world dam3d dam2d dam2d_1000 droplet3d droplet2d bubble2d shock3d 2>&1 | grep " grid " | cut -c1-260
