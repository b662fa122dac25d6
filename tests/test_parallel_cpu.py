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


def test_hole_filling_chain_equals_the_sequential_exchange_loop():
    """the algorithm of k_holefill (csrc/b200_comm.cuh) restated in Python against CommBrick::exchange's loop (comm_brick.cpp:628-650):
    same pack order, same final local indices, for random leaver sets (incl. leavers at the end, runs of leavers, everybody leaving)"""
    import numpy as np
    rng = np.random.default_rng(11)

    def reference(n, leave):
        atoms = list(range(n)); packed = []; nlocal = n; i = 0
        while i < nlocal:
            if leave[atoms[i]]:
                packed.append(atoms[i]); atoms[i] = atoms[nlocal - 1]; nlocal -= 1      # avec->copy(nlocal-1,i,1); nlocal--
            else:
                i += 1
        return packed, atoms[:nlocal]

    def chain(n, leave):
        a = [i for i in range(n) if leave[i]]                    # leaver indices ascending (x_list)
        m = len(a); index_of = list(range(n)); tail = n - 1; packed = []
        k = 0
        while k < m and a[k] <= tail:
            i = a[k]; cur = i
            while True:
                packed.append(cur)
                if i == tail:
                    tail -= 1; break
                t = tail; tail -= 1
                if leave[t]:
                    cur = t; continue
                index_of[t] = i
                break
            k += 1
        stay = [None] * (n - m)
        for atom in range(n):
            if not leave[atom]:
                stay[index_of[atom]] = atom
        return packed, stay

    for trial in range(400):
        n = int(rng.integers(1, 40))
        p = rng.choice([0.0, 0.1, 0.5, 0.9, 1.0])
        leave = [bool(v) for v in (rng.random(n) < p)]
        assert reference(n, leave) == chain(n, leave), (n, leave)
