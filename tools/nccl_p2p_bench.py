"""NCCL send/recv bandwidth between neighbouring ranks, as the halo uses it (one grouped call, both directions):
   torchrun --nproc-per-node N tools/nccl_p2p_bench.py"""
import os, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
right, left = (rank + 1) % world, (rank - 1) % world
for mb in (0.25, 1, 4, 16, 64):
    n = int(mb * 1e6 / 8)
    s1, s2 = torch.ones(n, dtype=torch.float64, device="cuda"), torch.ones(n, dtype=torch.float64, device="cuda")
    r1, r2 = torch.empty_like(s1), torch.empty_like(s1)
    def go():
        ops = [dist.P2POp(dist.isend, s1, right), dist.P2POp(dist.irecv, r1, left), dist.P2POp(dist.isend, s2, left), dist.P2POp(dist.irecv, r2, right)]
        for w in dist.batch_isend_irecv(ops): w.wait()
    for _ in range(3): go()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if rank == 0:
        print("message %6.2f MB x (2 sends + 2 recvs): %.3f ms -> %.1f GB/s sent per rank (per direction %.1f)" % (mb, ms, 2 * mb / ms, mb / ms), flush=True)
dist.destroy_process_group()
