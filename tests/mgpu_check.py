"""Multi-GPU correctness run (launched by torchrun, one rank per GPU); the comparison recipe is tests/mgpu_lib.py.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py [case ...]
  ... tests/mgpu_check.py --grid 2x2x2 dam3d droplet3d          (force a processor grid where the box allows it)
  ... tests/mgpu_check.py --empty-rank                           (one brick without atoms: the dry half of a dam break)
  ... tests/mgpu_check.py --balance x dam3d dam2d                (non-uniform bricks from parallel.balance_shift)
  ... tests/mgpu_check.py --vs-world shock3d                     (expected values from the P-rank oracle instead of the 1-rank fixture)
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mgpu_lib  # noqa: E402

DEFAULT = ["dam3d", "dam2d", "heat3d", "heat2d_rhosum", "droplet3d_static", "droplet2d_static", "droplet3d", "droplet2d_pcheat_skin",
           "droplet3d_heat", "bubble3d", "shock3d_shrink"]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    args = sys.argv[1:]
    grid = None
    if "--grid" in args:
        k = args.index("--grid"); grid = tuple(int(v) for v in args[k + 1].split("x")); del args[k:k + 2]
    balance = None
    if "--balance" in args:
        k = args.index("--balance"); balance = args[k + 1]; del args[k:k + 2]
    vs_world = "--vs-world" in args          # compare with the P-rank oracle even where the 1-rank fixture would do (decks whose result
    if vs_world:                             # depends on the local index order the ranks end up with, e.g. the two-type sph/idealgas deck)
        args.remove("--vs-world")
    failed = 0
    if "--empty-rank" in args:
        args.remove("--empty-rank")
        failed += empty_rank_case(rank, world, local)
        names = args
    else:
        names = args or DEFAULT
    for name in names:
        r = mgpu_lib.check_case(name, dist, rank, world, local, grid, balance=balance, vs_world=vs_world)
        if rank == 0:
            print(mgpu_lib.format_result(r), flush=True)
            failed += 0 if r["ok"] else 1
    flag = [failed]
    dist.broadcast_object_list(flag, src=0)
    dist.destroy_process_group()
    sys.exit(1 if flag[0] else 0)


def empty_rank_case(rank, world, local):
    """dam3d with the box stretched in x so that the last brick owns no atom at all: it still holds send lists / ghosts of its
    neighbour and must keep taking part in every halo and collective (hang in ncclSend/ncclRecv).  Expected values:
    the CPU oracle emulating the same ranks on the same stretched box."""
    import cases
    import harness
    from pworld import OracleWorld
    from util import relerr
    pkg = mgpu_lib.pkg
    case = cases.CASES["dam3d"]
    g = harness.load_golden("dam3d")
    (x0, y0, z0), (x1, y1, z1) = case.box
    import copy
    c2 = copy.copy(case)
    c2.box = ((x0, y0, z0), (x0 + (x1 - x0) * 2.2 * world / 2, y1, z1))      # atoms fill less than the first half of the box
    deck = c2.deck()
    grid = (world, 1, 1)
    brick = pkg.parallel.Brick(world, rank, deck.boxlo, deck.boxhi, 3, grid)
    st = harness.state_from(g, "init_", False)
    mine = brick.owns(st["x"])
    counts = [None] * world
    dist.all_gather_object(counts, int(mine.sum()))
    sim = pkg.B200Sim(deck, device=local, brick=brick, nccl_id=pkg.parallel.nccl_id(pkg.load(), dist))
    sim.set_atoms(**{k: np.ascontiguousarray(v[mine]) for k, v in st.items()})
    sim.setup(); sim.run(case.nsteps)
    out = sim.get_atoms(); sim.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    bad = 0
    if rank == 0:
        w = OracleWorld(c2.deck(), world, grid)
        w.set_atoms(**st); w.setup(); w.run(case.nsteps)
        want = w.get_atoms(); w.close()
        tags = np.concatenate([o["tag"] for o in gathered]); order = np.argsort(tags)
        errs = {k: relerr(np.concatenate([o[k] for o in gathered])[order], want[k]) for k in ("x", "v", "f", "rho", "e", "de", "drho")}
        ok = min(counts) == 0 and np.array_equal(tags[order], want["tag"]) and max(errs.values()) <= 1e-8
        print("%-24s grid %s atoms/rank %s (one brick empty) vs the oracle emulating the same ranks: %s  %s" % (
            "dam3d_empty_rank", grid, counts, "OK" if ok else "FAIL", {k: "%.1e" % v for k, v in errs.items()}), flush=True)
        bad = 0 if ok else 1
    dist.barrier()
    return bad


if __name__ == "__main__":
    main()
