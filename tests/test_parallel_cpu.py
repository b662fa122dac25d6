"""CPU, world_size 2 over gloo: the host-side decomposition logic of the multi-GPU path
(processor grid, sub-domains, neighbour ranks, ownership, object broadcast of the NCCL id slot)."""
import importlib
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

pkg = importlib.import_module("lammps-sph-multiphase_b200")
Brick = pkg.parallel.Brick


def test_grid_matches_surface_rule():
    assert pkg.parallel.proc_grid(8, (0, 0, 0), (1, 1, 1)) == (2, 2, 2)
    assert pkg.parallel.proc_grid(4, (0, 0, 0), (4, 1, 1)) == (4, 1, 1)
    assert pkg.parallel.proc_grid(4, (0, 0, 0), (1, 1, 0.01), dim=2)[2] == 1
    assert pkg.parallel.proc_grid(2, (0, 0, 0), (1.66, 1.12, 1.17)) == (2, 1, 1)


def test_bricks_tile_the_box():
    lo, hi = (0.0, -1.0, 0.5), (1.0, 2.0, 3.5)
    rng = np.random.default_rng(1)
    x = rng.uniform(lo, hi, size=(5000, 3))
    for world in (1, 2, 4, 8, 6):
        own = np.zeros(len(x), int)
        for r in range(world):
            b = Brick(world, r, lo, hi)
            own += b.owns(x)
            for d in range(3):       # my right neighbour's left neighbour is me
                right = Brick(world, b.procneigh[2 * d + 1], lo, hi)
                assert right.procneigh[2 * d] == r
                if b.myloc[d] < b.grid[d] - 1:
                    assert right.sublo[d] == b.subhi[d]     # bit-identical faces
        assert (own == 1).all()


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    payload = [b"id-from-rank0" if rank == 0 else None]
    dist.broadcast_object_list(payload, src=0)          # the path parallel.nccl_id uses
    b = Brick(world, rank, (0, 0, 0), (2.0, 1.0, 1.0))
    x = np.stack(np.meshgrid(np.arange(20) / 10 + 0.05, np.arange(10) / 10 + 0.05, np.arange(10) / 10 + 0.05, indexing="ij"), -1).reshape(-1, 3)
    mine = int(b.owns(x).sum())
    out = [None] * world
    dist.all_gather_object(out, (payload[0], mine, b.grid))
    if rank == 0:
        q.put(out)
    dist.destroy_process_group()


def test_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [o[0] for o in out] == [b"id-from-rank0"] * 2
    assert sum(o[1] for o in out) == 2000 and out[0][1] == 1000
    assert out[0][2] == (2, 1, 1)
