"""developer benchmark of the other BASELINE configs (not the contract bench): python tests/dev_bench.py c3 100 50
  c3 <nx> <steps>: periodic two-phase box (square_to_sphere deck), nx^3 particles, 4 multiphase pair styles, rebuild every step
  c4 <nx> <steps>: + sph/heatconduction/phasechange and fix phase_change
under torchrun (one rank per GPU) the box is split into LAMMPS-style bricks with NCCL halos (the C5 configuration):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/dev_bench.py c3 252 20"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
pkg = importlib.import_module("lammps-sph-multiphase_b200")


def lattice_atoms(nx, two_phase_cube=0.2, jitter=0.0):
    dx = 1.0 / nx
    i = np.arange(nx) * dx
    x = np.stack(np.meshgrid(i, i, i, indexing="ij"), -1).reshape(-1, 3)
    if jitter:
        x = x + np.random.default_rng(1).uniform(-jitter * dx, jitter * dx, x.shape)
        x %= 1.0
    n = len(x)
    typ = np.where((np.abs(x - 0.5) <= two_phase_cube).all(1), 2, 1).astype(np.int32)
    return dict(x=x, v=np.zeros((n, 3)), rho=np.ones(n), e=np.where(typ == 1, 1.0, 1.5), cv=np.where(typ == 1, 1.0, 2.0),
                rmass=np.full(n, dx ** 3), type=typ, mask=np.ones(n, np.int32), tag=np.arange(1, n + 1, dtype=np.int32))


def main():
    kind, nx, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    if kind == "c3":
        case = cases._droplet("c3", 3, nx, steps)
        atoms = lattice_atoms(nx)
    else:
        case = cases._bubble("c4", 3, nx, steps)
        atoms = lattice_atoms(nx, 0.12, jitter=0.2)
        dx = 1.0 / nx
        atoms["rho"] = np.where(atoms["type"] == 2, 0.1, 1.0); atoms["rmass"] = atoms["rho"] * dx ** 3
        atoms["cv"] = np.where(atoms["type"] == 2, 0.06, 0.04); atoms["e"] = np.where(atoms["type"] == 2, 0.06 * 0.6, 0.04)
        atoms["mask"] = np.where(atoms["type"] == 2, 3, 1).astype(np.int32)
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    deck = case.deck()
    brick = nid = dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("gloo")
        brick = pkg.parallel.Brick(world, rank, deck.boxlo, deck.boxhi, 3)
        nid = pkg.parallel.nccl_id(pkg.load(), dist)
        mine = brick.owns(atoms["x"])
        atoms = {k: np.ascontiguousarray(v[mine]) for k, v in atoms.items()}
    sim = pkg.B200Sim(deck, device=local, brick=brick, nccl_id=nid)
    sim.set_atoms(**atoms)
    sim.setup(); sim.run(5); sim.sync()
    sim.set_timing(True)
    if dist: dist.barrier()
    t0 = time.perf_counter(); sim.run(steps); sim.sync()
    if dist: dist.barrier()
    t = time.perf_counter() - t0
    n = sim.natoms()
    if dist:          # whole-job figures + Newton's third law over the whole periodic box
        f = sim.get_atoms()["f"]
        tot = [None] * world
        dist.all_gather_object(tot, (n[0], n[1], f.sum(0), np.abs(f).sum(0)))
        if rank == 0:
            ntot = sum(v[0] for v in tot); fs = sum(v[2] for v in tot); fa = sum(v[3] for v in tot)
            print("%s nx=%d ranks=%d grid=%s particles=%d (per rank %s, ghosts %s)  %.3f ms/step  %.1f M particle-steps/s  sum f / sum |f| = %.1e" % (
                kind, nx, world, brick.grid, ntot, [v[0] for v in tot], [v[1] for v in tot], 1e3 * t / steps, ntot * steps / t / 1e6,
                np.abs(fs).max() / fa.max()))
        if rank != 0:
            sim.close(); dist.barrier(); dist.destroy_process_group(); return
    print("%s nx=%d particles=%d ghosts=%d  %.3f ms/step  %.1f M particle-steps/s  counters %s" % (kind, nx, n[0], n[1], 1e3 * t / steps, n[0] * steps / t / 1e6, sim.counters()))
    for k, (ms, calls) in sim.timers().items():
        if calls:
            print("   %-24s %9.3f ms total  %8.3f ms/call  %5.1f %%" % (k, ms, ms / calls, 100 * ms / (1e3 * t)))
    if dist:
        sim.close(); dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
