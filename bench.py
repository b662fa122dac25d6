#!/usr/bin/env python
"""bench.py -- particle-steps/s of the SPH hot path (BASELINE.json metric) on N B200s.

Workload (N=1): BASELINE.json configs[1], "3D dam break, sph/taitwater + sph/rhosum, 1M particles on
1xB200", built as SURVEY.md 8(d) C2 specifies: sc lattice dx = 0.01, water block 100x100x80 in the
corner of an open-top tank (inner 160x106x106 sites, 3-layer walls), h = 3 dx, c = 30,
`pair_style hybrid/overlay sph/rhosum 1 sph/taitwater`, gravity on the water, `fix meso` (water) and
`fix meso/stationary` (walls), `neigh_modify every 5 delay 0 check no`, skin 0.3 h, dt = 1e-4.
A "step" is one Verlet timestep of that deck (neighbor rebuild every 5th step, density summation,
Tait force + viscosity, gravity, velocity-Verlet update of x, v, rho, e).

  python bench.py --gpus N --steps K --warmup W          # the CUDA engine through the C-ABI
  python bench.py --impl reference --steps K --warmup W  # the reference's own CPU path (oracle/_ref)

For N > 1 the tank is tiled N times along x (one C2-sized water column per GPU) and decomposed into
N bricks: ghost exchange, reverse accumulation and atom migration run over NCCL (weak scaling).

Besides the headline (`value`, C2) every line carries
  "configs": C3 (two-phase droplet box, nx = 160: 4.1 M particles, 1 GPU) at N = 1; C5_strong (the same periodic multiphase box at
             nx = 400: 64 M particles, brick grid from ProcMap::onelevel_grid, SAME total at every N incl. N = 1 -> strong scaling);
             C4 (16 M particles, + heat/phasechange and fix phase_change) at N = 2, 4 -- each with ms_per_step, stage_ms,
             particle_steps_s, roofline (228 / 572 B per particle-step) and sum f / sum |f| over the periodic box;
  "parity":  at N > 1, AFTER the timed regions: the fixtures dam3d, heat3d, droplet3d_static on the N ranks against the 1-rank
             reference fixtures and droplet3d, droplet3d_heat against the CPU oracle emulating the same brick grid (2x2x2 at
             N = 8) -- tests/mgpu_lib.py, checker role only.  The run exits non-zero if parity is not ok.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions.
"""
import argparse
import importlib
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL prints its version banner on stdout when NCCL_DEBUG=VERSION; stdout carries exactly one JSON line here
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

METRIC = "particle-steps/sec (SPH rho+force+heat loop)"
UNIT = "particle-steps/s"
# algorithmic bytes per particle per launch of the dominant kernel (single-phase):
#   sph/taitwater stage: R x24 + vest24 + rho8 + type4, W f24 + drho8 + de8 = 100 B
BYTES_FORCE = 100
FP64_NOTE = "profiles/r02_final_full.txt: k_tile_force<K_TAIT> under ncu --set full (1 028 768 particles, 0.459 ms): sm__pipe_fp64_cycles_active 50.4 %, shared-memory data pipe 69.5 %, 47 % of the fp64 lanes (dadd + dfma + dmul thread-instructions); 31 fp64 instructions per list entry (round 1: 42, 50.6 % at 0.539 ms)"
BYTES_STEP = 464          # whole single-phase step (rhosum 36 + taitwater 100 + fix meso 200 + 128)
BYTES_MP_LOOP, BYTES_MP_STEP = 228, 572        # multiphase density + colorgradient + force loop / whole step
BYTES_C4_LOOP, BYTES_C4_STEP = 252, 596        # + heat/phasechange in the fused force pass
FP64_LANES_PER_SM = 64    # B200: 64 fp64 FMA lanes per SM and clock
PARITY_CASES = ("dam3d", "heat3d", "droplet3d_static", "droplet3d", "droplet3d_heat")


# ----------------------------------------------------------------------------- workload ----
def dam_break_3d(scale=1.0, tiles=1):
    """C2 geometry; scale shrinks every edge (scale=0.5 -> ~1/8 of the particles).
    tiles > 1 (weak scaling): `tiles` complete tanks (walls included) side by side in x, one water column each,
    so an (N,1,1) brick decomposition gives every GPU exactly the C2 load of the 1-GPU run (1 028 768 particles at scale 1)."""
    dx = 0.01
    nwx, nwy, nwz = (max(4, int(round(v * scale))) for v in (100, 100, 80))      # water block (sites)
    nix, niy, niz = (max(6, int(round(v * scale))) for v in (160, 106, 106))     # tank inner size
    nl = 3                                                                        # wall layers
    unit = nix + 2 * nl
    NX, NY, NZ = unit * tiles, niy + 2 * nl, niz + nl                             # open top
    ix, iy, iz = np.meshgrid(np.arange(NX), np.arange(NY), np.arange(NZ), indexing="ij")
    ix, iy, iz = ix.ravel(), iy.ravel(), iz.ravel()
    wall = ((ix % unit) < nl) | ((ix % unit) >= unit - nl) | (iy < nl) | (iy >= NY - nl) | (iz < nl)      # every unit is a complete tank: the same particle count per GPU at every N
    water = (~wall) & ((ix % unit) >= nl) & ((ix % unit) < nl + nwx) & (iy < nl + nwy) & (iz < nl + nwz)
    keep = wall | water
    x = np.stack([(ix[keep] + 0.5) * dx, (iy[keep] + 0.5) * dx, (iz[keep] + 0.5) * dx], axis=1)
    typ = np.where(wall[keep], 2, 1).astype(np.int32)
    n = len(typ)
    box = ((0.0, 0.0, 0.0), (NX * dx, NY * dx, (NZ + 8) * dx))
    h, c = 3 * dx, 30.0
    atoms = dict(x=np.ascontiguousarray(x), v=np.zeros((n, 3)), rho=np.full(n, 1000.0), e=np.zeros(n), cv=np.ones(n),
                 type=typ, mask=np.where(typ == 2, 1 | 2, 1 | 4).astype(np.int32), tag=np.arange(1, n + 1, dtype=np.int32))
    params = dict(dx=dx, h=h, c=c, m=1000.0 * dx ** 3, skin=0.3 * h, dt=1.0e-4, box=box)   # dt = 0.1 h / c
    return atoms, params


def make_deck(pkg, params):
    d = pkg.Deck(dimension=3, boundary="f f f", box=params["box"], atom_style="meso", ntypes=2, units="si")
    d.group("bc"); d.group("water")                      # bits 2, 4 (group bc type 2; group water type 1)
    d.mass("*", params["m"])
    d.pair_style("hybrid/overlay", "sph/rhosum 1", "sph/taitwater")
    d.pair_coeff("* *", "sph/taitwater", 1000.0, params["c"], 1.0, params["h"])
    d.pair_coeff("1 1", "sph/rhosum", params["h"])
    d.fix("gfix", "water", "gravity", -9.81, "vector", 0, 0, 1)
    d.fix("integrate_water", "water", "meso")
    d.fix("integrate_bc", "bc", "meso/stationary")
    d.neigh_modify(every=5, delay=0, check="no")
    d.neighbor(params["skin"])
    d.timestep(params["dt"])
    return d.init()


def write_lammps_case(dirname, atoms, params, nsteps, warmup):
    """the same deck in the reference's language + a data file with the same particles (%.17g)"""
    n = len(atoms["type"])
    (x0, y0, z0), (x1, y1, z1) = params["box"]
    with open(os.path.join(dirname, "data.c2"), "w") as f:
        f.write("C2 3-D dam break (generated by bench.py)\n\n%d atoms\n2 atom types\n\n" % n)
        f.write("%.17g %.17g xlo xhi\n%.17g %.17g ylo yhi\n%.17g %.17g zlo zhi\n\nAtoms\n\n" % (x0, x1, y0, y1, z0, z1))
        x = atoms["x"]; t = atoms["type"]
        lines = ["%d %d 1000.0 0.0 1.0 %.17g %.17g %.17g\n" % (i + 1, t[i], x[i, 0], x[i, 1], x[i, 2]) for i in range(n)]
        f.writelines(lines)
    deck = """units si
dimension 3
boundary f f f
newton on
atom_style meso
read_data data.c2
mass * %(m).17g
group bc type 2
group water type 1
pair_style hybrid/overlay sph/rhosum 1 sph/taitwater
pair_coeff * * sph/taitwater 1000.0 %(c).17g 1.0 %(h).17g
pair_coeff 1 1 sph/rhosum %(h).17g
fix gfix water gravity -9.81 vector 0 0 1
fix integrate_water water meso
fix integrate_bc bc meso/stationary
neigh_modify every 5 delay 0 check no
neighbor %(skin).17g bin
timestep %(dt).17g
thermo 1000
""" % params
    if warmup:
        deck += "run %d\n" % warmup
    deck += "run %d\n" % nsteps
    with open(os.path.join(dirname, "in.c2"), "w") as f:
        f.write(deck)


def run_reference(atoms, params, nsteps, warmup, replicas):
    """time the reference's own CPU path: `replicas` concurrent lmp_serial processes (no MPI in this image),
    aggregate particle-steps/s from each process' own `Loop time` line of the LAST run (finish.cpp:100-125)"""
    exe = os.path.join(ROOT, "oracle", "_ref", "lmp_serial")
    if not os.path.exists(exe):
        return None
    tmp = tempfile.mkdtemp(prefix="b200ref_")
    try:
        write_lammps_case(tmp, atoms, params, nsteps, warmup)
        t0 = time.time()
        procs = [subprocess.Popen([exe, "-in", "in.c2", "-log", "none", "-screen", "out.%d" % r], cwd=tmp,
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for r in range(replicas)]
        for p in procs:
            p.wait()
        wall = time.time() - t0
        rate, split = 0.0, None
        for r in range(replicas):
            txt = open(os.path.join(tmp, "out.%d" % r)).read()
            loops = re.findall(r"Loop time of ([0-9.eE+-]+) on \d+ procs for (\d+) steps with (\d+) atoms", txt)
            if not loops:
                raise RuntimeError("reference run failed:\n" + txt[-2000:])
            t, s, n = loops[-1]
            rate += int(n) * int(s) / float(t)
            if split is None:
                split = {k: float(v) for k, v in re.findall(r"(Pair|Neigh|Comm|Outpt|Other)\s+time \(%\) = [0-9.eE+-]+ \(([0-9.]+)\)", txt)[-5:]}
        return dict(value=rate, wall_s=wall, split=split)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)



# ------------------------------------------------------------- multiphase configs ----
def lattice_brick(nx, brick, cube=0.2, jitter=0.0, c4=False):
    """this rank's part of the nx^3 simple-cubic lattice of the unit periodic box (the square_to_sphere / bubble decks at scale):
    type 2 inside the centred cube.  Per-dimension ownership tests on the 1-D coordinates, so the ranks' parts tile the box
    exactly; tags are the global lattice index + 1; the jitter is a hash of the tag (the same on any decomposition)."""
    dx = 1.0 / nx
    c1 = np.arange(nx) * dx
    sel = []
    for d in range(3):
        lo, hi = (brick.sublo[d], brick.subhi[d]) if brick is not None else (0.0, 1.0)
        sel.append(np.nonzero((c1 >= lo) & (c1 < hi))[0])
    ix, iy, iz = np.meshgrid(sel[0], sel[1], sel[2], indexing="ij")
    ix, iy, iz = ix.ravel(), iy.ravel(), iz.ravel()
    n = len(ix)
    gid = (ix.astype(np.int64) * nx + iy) * nx + iz
    x = np.stack([c1[ix], c1[iy], c1[iz]], axis=1)
    if jitter:
        for d in range(3):
            u = ((gid * np.int64(2654435761) + np.int64(40503 * (d + 1))) % np.int64(1 << 20)).astype(np.float64) / float(1 << 20)
            x[:, d] += (2.0 * u - 1.0) * jitter * dx
        x %= 1.0
        if brick is not None:      # a jittered atom may have left the brick by a hair: put it back (ownership stays a partition)
            for d in range(3):
                x[:, d] = np.clip(x[:, d], brick.sublo[d], np.nextafter(brick.subhi[d], -1.0))
    typ = np.where((np.abs(x - 0.5) <= cube).all(1), 2, 1).astype(np.int32)
    a = dict(x=np.ascontiguousarray(x), v=np.zeros((n, 3)), rho=np.ones(n), e=np.where(typ == 1, 1.0, 1.5), cv=np.where(typ == 1, 1.0, 2.0),
             rmass=np.full(n, dx ** 3), type=typ, mask=np.ones(n, np.int32), tag=(gid + 1).astype(np.int32))
    if c4:
        a["rho"] = np.where(typ == 2, 0.1, 1.0); a["rmass"] = a["rho"] * dx ** 3
        a["cv"] = np.where(typ == 2, 0.06, 0.04); a["e"] = np.where(typ == 2, 0.06 * 0.6, 0.04)
        a["mask"] = np.where(typ == 2, 3, 1).astype(np.int32)
    return a


def run_mp_config(pkg, kind, nx, steps, warmup, rank, world, local, dist, torch, peak):
    """one multiphase config: the tests' own deck generators (tests/cases.py: _droplet = examples/USER/sph/square_to_sphere,
    _bubble = bubble_random with fix phase_change) at nx^3 particles on `world` bricks; device-timed K steps, max over ranks"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    c4 = kind == "C4"
    case = cases._bubble("c4", 3, nx, steps) if c4 else cases._droplet("c3", 3, nx, steps)
    deck = case.deck()
    brick = nid = None
    if world > 1:
        brick = pkg.parallel.Brick(world, rank, deck.boxlo, deck.boxhi, 3)
        nid = pkg.parallel.nccl_id(pkg.load(), dist)
    t0 = time.perf_counter()
    atoms = lattice_brick(nx, brick, 0.12 if c4 else 0.2, 0.2 if c4 else 0.0, c4)
    n = len(atoms["type"])
    sim = pkg.B200Sim(deck, device=local, brick=brick, nccl_id=nid)
    sim.set_atoms(**atoms)
    del atoms
    sim.setup()
    sim.run(max(warmup, 3))
    sim.sync()
    t_setup = time.perf_counter() - t0
    sim.set_timing(True)
    c0 = sim.counters()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    sim.run(steps)
    ev1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    nl, ng = sim.natoms()
    f = sim.get_atoms(("f",))["f"]
    red = torch.tensor([n, nl, ng] + list(f.sum(0)) + list(np.abs(f).sum(0)), dtype=torch.float64, device="cuda")
    mx = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(red); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    red = red.cpu().numpy(); ms = float(mx.item())
    c1 = sim.counters(); timers = sim.timers()
    sim.close()
    ntot = int(red[0])
    rate = ntot * steps / (ms * 1e-3)
    loop_ms = sum(timers.get(k, (0.0, 0))[0] for k in ("density", "colorgradient", "records", "force"))
    bl, bs = (BYTES_C4_LOOP, BYTES_C4_STEP) if c4 else (BYTES_MP_LOOP, BYTES_MP_STEP)
    out = {"workload": ("C4: evaporating two-phase box, 5 multiphase sub-styles incl. heatconduction/phasechange + fix phase_change" if c4 else
                        "periodic two-phase box (square_to_sphere deck): rhosum/multiphase + colorgradient + taitwater/multiphase + surfacetension, rebuilt every step"),
           "nx": nx, "particles_total": ntot, "particles_end": int(red[1]), "ghosts_total": int(red[2]), "n_gpus": world,
           "grid": list(brick.grid) if brick is not None else [1, 1, 1], "steps": steps, "warmup": max(warmup, 3),
           "ms_per_step": ms / steps, "particle_steps_s": rate, "stage_ms": {k: round(v[0], 3) for k, v in timers.items() if v[1]},
           "neighbor_builds": c1["builds"] - c0["builds"], "gpu_launches": c1["launches"] - c0["launches"], "inserted": c1["inserted"],
           "sum_f_over_sum_abs_f": float(np.abs(red[3:6]).max() / max(red[6:9].max(), 1e-300)),
           "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak, "bytes_per_particle_loop": bl, "bytes_per_particle_step": bs,
                        "step_GBs": bs * rate / 1e9, "step_frac": bs * rate / 1e9 / (peak * world),
                        "loop_GBs": (bl * ntot * steps / (loop_ms * 1e-3) / 1e9) if loop_ms else None,
                        "loop_frac": (bl * ntot * steps / (loop_ms * 1e-3) / 1e9 / (peak * world)) if loop_ms else None,
                        "note": "loop = density + colorgradient + records + force stage timers of rank 0 (per-rank work, whole-job bytes)"},
           "setup_seconds": round(t_setup, 2)}
    return out


def run_dam_balance(pkg, scale, steps, warmup, rank, world, local, dist, torch):
    """ONE dam-break tank (C2 geometry at `scale`, 8 M particles at scale 2) on `world` bricks -- the water column sits in
    one corner, so uniform bricks are badly loaded.  Measured twice: uniform cuts, and the cuts of `balance 1.05 shift xyz 10 1.05`
    (parallel.balance_shift = Balance::shift, src/balance.cpp:632-790)."""
    atoms, params = dam_break_3d(scale)
    deck = make_deck(pkg, params)
    grid = pkg.parallel.proc_grid(world, deck.boxlo, deck.boxhi, 3)
    out = {"workload": "C2 geometry at edge scale %g, ONE tank (not tiled), %d particles, %s bricks" % (scale, len(atoms["type"]), "x".join(map(str, grid))),
           "particles_total": len(atoms["type"]), "grid": list(grid), "steps": steps}
    for label in ("uniform", "balanced"):
        splits = pkg.parallel.balance_shift(atoms["x"], deck.boxlo, deck.boxhi, grid, "xyz", 10, 1.05) if label == "balanced" else None
        brick = pkg.parallel.Brick(world, rank, deck.boxlo, deck.boxhi, 3, grid, splits)
        mine = brick.owns(atoms["x"])
        sim = pkg.B200Sim(deck, device=local, brick=brick, nccl_id=pkg.parallel.nccl_id(pkg.load(), dist))
        sim.set_atoms(**{k: np.ascontiguousarray(v[mine]) for k, v in atoms.items()})
        sim.setup(); sim.run(max(warmup, 3)); sim.sync()
        dist.barrier(); torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(); sim.run(steps); ev1.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([ev0.elapsed_time(ev1)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        counts = [None] * world
        dist.all_gather_object(counts, int(sim.natoms()[0]))
        sim.close()
        ms = float(t.item()) / steps
        out[label] = {"ms_per_step": ms, "particle_steps_s": len(atoms["type"]) / (ms * 1e-3), "atoms_per_rank": counts,
                      "imbalance_max_over_mean": max(counts) / (sum(counts) / world),
                      "splits": None if splits is None else [[round(float(v), 6) for v in sp] for sp in splits]}
    out["speedup_balanced_over_uniform"] = out["uniform"]["ms_per_step"] / out["balanced"]["ms_per_step"]
    out["ms_per_step"] = out["balanced"]["ms_per_step"]; out["particle_steps_s"] = out["balanced"]["particle_steps_s"]      # the record's own figures: the balanced run
    return out


def run_parity(dist, rank, world, local):
    """fixtures on the N ranks (tests/mgpu_lib.py); 2x2x2 bricks at N = 8"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import mgpu_lib
    prefer = {8: (2, 2, 2), 4: (2, 2, 1), 2: (2, 1, 1)}.get(world)
    cases_out, ok, worst = [], True, 0.0
    for name in PARITY_CASES:
        r = mgpu_lib.check_case(name, dist, rank, world, local, prefer)
        if rank == 0:
            cases_out.append({k: r.get(k) for k in ("name", "grid", "against", "ok", "err", "err_elem")})
            ok = ok and bool(r["ok"]); worst = max(worst, float(r.get("err", 0.0)))
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    return {"ok": bool(flag[0]), "max_err": worst, "grid": list(prefer) if prefer else None, "cases": cases_out,
            "tolerance": "10x the case's trajectory tolerance (1e-8) vs fixtures, 100x (1e-7) vs the P-rank oracle; norm-wise max|a-b|/max|b|, element-wise figure alongside"}

# ------------------------------------------------------------------------------- clocks ----
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.p, self.path, self.skip = gpu, None, None, 0

    def start(self):
        if not shutil.which("nvidia-smi"):
            return
        fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
        os.close(fd)
        self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                  stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)

    def ready(self, timeout=5.0):
        """wait for the sampler's first line: nvidia-smi has attached to the devices and only polls from here on"""
        t0 = time.time()
        while self.p and time.time() - t0 < timeout and os.path.getsize(self.path) == 0:
            time.sleep(0.02)

    def mark(self):
        """the timed region starts here: samples written so far (nvidia-smi's start-up, the warm-up) do not count"""
        if self.p:
            self.skip = sum(1 for _ in open(self.path))

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.p:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        lines = open(self.path).read().splitlines()
        lines = lines[self.skip:] or lines[-1:]          # a region shorter than one sampling period: the sample just before it
        for line in lines:
            c = [v.strip() for v in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


# -------------------------------------------------------------------------------- main ----
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="edge scale of the C2 geometry (1.0 = 1M particles)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C3 / C4 / C5_strong sub-records")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-GPU parity block (N > 1)")
    ap.add_argument("--c3-nx", type=int, default=160)
    ap.add_argument("--c4-nx", type=int, default=252)
    ap.add_argument("--c5-nx", type=int, default=400)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    pkg = importlib.import_module("lammps-sph-multiphase_b200")

    if args.impl == "reference":
        if rank != 0:
            return
        # bounded sample: the C2 geometry at half edge length (~1/8 of the particles), same deck, all host cores
        cores = os.cpu_count() or 1
        replicas = max(1, min(cores, 64))
        atoms, params = dam_break_3d(0.5 * args.scale)
        n = len(atoms["type"])
        t0 = time.time()
        r = run_reference(atoms, params, max(1, args.steps), max(0, args.warmup), replicas)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/lmp_serial not built (run python __graft_entry__.py build where /root/reference exists)"}))
            return
        sample = "C2 dam-break deck at 0.5 edge scale (%d particles), %d timed steps after %d warm-up, %d concurrent lmp_serial replicas (no MPI in image); Loop time per replica" % (n, args.steps, args.warmup, replicas)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * n * replicas / r["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": {"workload": "C2 3D dam break, sph/rhosum + sph/taitwater (bounded sample)", "particles": n, "replicas": replicas},
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": replicas, "kind": "reference", "sample": sample, "split_pct": r["split"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "wall_s": time.time() - t0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the SPH hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    atoms, params = dam_break_3d(args.scale, tiles=world)
    deck = make_deck(pkg, params)
    brick, nid = None, None
    if world > 1:      # LAMMPS-style brick decomposition; halo exchange and migration run over NCCL inside the library
        brick = pkg.parallel.Brick(world, rank, deck.boxlo, deck.boxhi, 3)
        nid = pkg.parallel.nccl_id(pkg.load(), dist)
        mine = brick.owns(atoms["x"])
        atoms = {k: np.ascontiguousarray(v[mine]) for k, v in atoms.items()}
    n = len(atoms["type"])
    ntot = n
    if world > 1:
        t = torch.tensor([n], device="cuda", dtype=torch.int64); dist.all_reduce(t); ntot = int(t.item())
    sim = pkg.B200Sim(deck, device=local, brick=brick, nccl_id=nid)
    sim.set_atoms(**atoms)
    sim.setup()
    # one sampler per job (concurrent nvidia-smi pollers contend for the driver lock and stall the ranks' API calls), started before the
    # warm-up so that its start-up (device attach, ~1 s) is over when the timed region begins; only samples taken inside the region count
    clocks = ClockSampler(local)
    if rank == 0 and not os.environ.get("BENCH_NO_CLOCKS"):
        clocks.start()
    sim.run(max(args.warmup, 3))
    sim.sync()

    # ---- device-resident timed region: K Verlet steps, CUDA events on the launching (default) stream ----
    sim.set_timing(True)
    c0 = sim.counters()
    clocks.ready()
    barrier()
    clocks.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    sim.run(args.steps)
    ev1.record()
    barrier()
    clk = clocks.stop()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    c1 = sim.counters()
    timers = sim.timers()
    sim.set_timing(False)
    value = ntot * args.steps / (ms * 1e-3)

    # ---- end to end through the C-ABI with HOST buffers: pinned host arrays -> b200_set_atoms ->
    #      b200_setup -> b200_run(K) -> b200_get_atoms -> host arrays (one deck-level `run K`) ----
    e2e = None
    if not args.no_e2e:
        pinned = {}
        for k, v in atoms.items():
            t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
            pinned[k] = t.numpy()
        out_names = ("x", "v", "vest", "f", "rho", "drho", "e", "de")
        h2d = sum(v.nbytes for v in pinned.values())
        nid2 = pkg.parallel.nccl_id(pkg.load(), dist) if world > 1 else None
        sim2 = pkg.B200Sim(deck, device=local, brick=brick, nccl_id=nid2)
        # result buffers are the caller's pinned arrays too (the LAMMPS shell hands the engine its own atom arrays);
        # sized with head room for atoms that migrate in (N > 1)
        cap = int(n * 1.25) + 1024
        outbuf = {k: torch.empty((cap, 3) if k in ("x", "v", "vest", "f") else (cap,), dtype=torch.float64).pin_memory().numpy() for k in out_names}
        sim2.set_atoms(**pinned); sim2.setup(); sim2.run(2); sim2.get_atoms(out_names, out=outbuf); sim2.sync()   # warm allocations
        barrier()
        t0 = time.perf_counter()
        sim2.set_atoms(**pinned)
        sim2.setup()
        sim2.run(args.steps)
        res = sim2.get_atoms(out_names, out=outbuf)
        sim2.sync()
        t_e2e = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([t_e2e], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); t_e2e = float(t.item())
        d2h = sum(v.nbytes for v in res.values())
        # where the end-to-end time goes: the same sequence once more with a stream synchronisation after every call (untimed above)
        parts = {}
        for nm, fn in (("set_atoms", lambda: sim2.set_atoms(**pinned)), ("setup", sim2.setup), ("run", lambda: sim2.run(args.steps)),
                       ("get_atoms", lambda: sim2.get_atoms(out_names, out=outbuf))):
            t1 = time.perf_counter(); fn(); sim2.sync(); parts[nm] = round(time.perf_counter() - t1, 5)
        e2e = {"value": ntot * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
               "mode": "pinned host arrays -> b200_set_atoms + b200_setup + b200_run(K) + b200_get_atoms -> pinned host arrays (one host round trip per K-step run, setup included)",
               "seconds": t_e2e, "breakdown_s": parts}
        sim2.close()

    # ---- roofline of the dominant kernel (fused force pass, k_force<K_TAIT>) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0)); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    f_ms, f_calls = timers.get("force", (0.0, 0))
    roof = None
    if f_calls:
        dur = f_ms / f_calls * 1e-3
        achieved = BYTES_FORCE * n / dur / 1e9
        traffic = None
        for tp in ("r02_force_fp64.json", "r01_force_dram_bytes.json"):      # newest ncu record of this kernel first
            tp = os.path.join(ROOT, "profiles", tp)
            if traffic is None and os.path.exists(tp):
                try:
                    traffic = json.load(open(tp)).get("dram_bytes_per_launch")
                except Exception:
                    pass
        roof = {"bound": "hbm", "kernel": "k_tile_force<K_TAIT> (sph/taitwater pass, shared-memory tile path)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": BYTES_FORCE * n, "avg_launch_ms": dur * 1e3,
                "share_of_step": f_ms / ms, "whole_step_GBs": BYTES_STEP * n * args.steps / (ms * 1e-3) / 1e9,
                "fp64_pipe": FP64_NOTE,
                "note": "bound by the fp64 pipe and the shared-memory data pipe together, not by HBM, at ~110 list entries per particle (SURVEY 8d): the HBM fraction cannot approach 1; "
                        "fp64_frac = fp64 thread-instructions per launch (ncu) / live launch time / fp64 lane peak"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2: 3D dam break, hybrid/overlay sph/rhosum 1 + sph/taitwater, fix gravity/meso/meso-stationary, %d particles per GPU" % n,
                       "particles_per_gpu": n, "neighbors_per_particle_list": c1["max_neighbors"], "rebuild_every": 5,
                       "parallelism": "1 GPU" if world == 1 else "%s brick decomposition, NCCL halo exchange + migration, tank tiled %dx in x (weak scaling)" % ("x".join(str(v) for v in brick.grid), world),
                       "particles_total": ntot,
                       "l2": "working set (particle state + neighbor rows, %.0f MB) exceeds the 126 MB L2; no explicit flush" % ((n * 200 + n * c1["row_stride"] * 4) / 1e6)},
            "clocks": clk, "gpu_launches": c1["launches"] - c0["launches"], "neighbor_builds": c1["builds"] - c0["builds"],
            "stage_ms": {k: round(v[0], 3) for k, v in timers.items() if v[1]}, "e2e": e2e, "roofline": roof}

    # ---- fp64-pipe fraction of the dominant kernel: fp64 thread-instructions of one launch (ncu, profiles/r02_force_fp64.json:
    #      smsp__sass_thread_inst_executed_op_fp64_pred_on of the same kernel on the same workload) / launch time measured live
    #      / (SMs x 64 fp64 lanes x SM clock under load) ----
    if roof is not None:
        fp = os.path.join(ROOT, "profiles", "r02_force_fp64.json")
        sm_hz = 1e6 * float((clk or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
        nsm = torch.cuda.get_device_properties(local).multi_processor_count
        try:
            rec = json.load(open(fp))
            per_particle = float(rec["fp64_thread_inst_per_launch"]) / float(rec["particles"])
            roof["fp64_frac"] = per_particle * n / (roof["avg_launch_ms"] * 1e-3) / (nsm * FP64_LANES_PER_SM * sm_hz)
            roof["fp64_thread_inst_per_particle"] = per_particle
            roof["fp64_source"] = "profiles/r02_force_fp64.json (ncu count) x live launch time; peak = %d SMs x 64 lanes x %.0f MHz" % (nsm, sm_hz / 1e6)
        except Exception:
            roof["fp64_frac"] = None

    # ---- the other BASELINE configs as sub-records (device-timed the same way; C2 stays the headline) ----
    configs = {}
    if not args.no_configs:
        plan = []
        if world == 1:
            plan.append(("C3", "C3", args.c3_nx, 20, 5))
        if world in (1, 2, 4):
            plan.append(("C4", "C4", args.c4_nx, 10, 3))
        plan.append(("C5_strong", "C5", args.c5_nx, 5, 3))
        sim.close(); sim = None
        torch.cuda.empty_cache()
        if world > 1:
            configs["C2_one_tank_8M"] = run_dam_balance(pkg, 2.0, 20, 5, rank, world, local, dist, torch)
        for key, kind, nx, k, w in plan:
            try:
                configs[key] = run_mp_config(pkg, kind, nx, k, w, rank, world, local, dist, torch, peak)
            except Exception as e:          # e.g. out of memory on a smaller device: report, keep the headline
                configs[key] = {"error": str(e)[:300], "nx": nx}
                if world > 1:
                    raise
    line["configs"] = configs

    # ---- multi-GPU parity, driver-visible (N > 1): fixtures / P-rank oracle on the same ranks, after every timed region ----
    parity = None
    if world > 1 and not args.no_parity:
        if sim is not None:
            sim.close(); sim = None
        parity = run_parity(dist, rank, world, local)
    line["parity"] = parity

    if rank == 0 and not args.no_cpu_baseline and world == 1:
        # bounded CPU sample: same deck at 0.5 edge scale, 1 replica (1 core), ~10-30 s
        a2, p2 = dam_break_3d(0.5 * args.scale)
        r = run_reference(a2, p2, 20, 5, 1)
        if r is not None:
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": 1, "kind": "reference",
                                    "sample": "oracle/_ref/lmp_serial (unmodified reference), C2 deck at 0.5 edge scale (%d particles), 20 steps after 5 warm-up, 1 process" % len(a2["type"]),
                                    "split_pct": r["split"]}
        else:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    if rank == 0:
        print(json.dumps(line))
    if sim is not None:
        sim.close()
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        sys.exit(3)


if __name__ == "__main__":
    main()
