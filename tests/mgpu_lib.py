"""Multi-GPU parity of one case on the ranks of a running torch.distributed job (one rank per GPU): the case is decomposed into
bricks, run through the C-ABI with NCCL halo exchange / migration, gathered by tag on rank 0 and compared

  * with the reference fixture of the SAME deck (tests/golden: what the reference's CPU path produced on one rank) when the
    result does not depend on the decomposition (single-phase decks, static multiphase decks);
  * with the CPU oracle emulating the same P ranks (tests/pworld.py) when it does: moving multiphase decks read one-step-stale
    ghost rho / colorgradient, in the reference too.  The P-rank emulation has no reference MPI run behind it (no
    MPI in this image); it is pinned on decomposition-independent decks only (tests/test_world_cpu.py);
  * fix phase_change decks draw one RNG stream per rank (fix_phase_change.cpp:107): they too are compared with the P-rank oracle, which
    walks one stream per emulated rank and numbers the new atoms as Atom::tag_extend does (sph_oracle.c w_phase_change).

Used by tests/mgpu_check.py (pytest -m gpu on >= 2 GPUs) and by bench.py's parity block at N > 1 (checker role only, after the
timed region)."""
import importlib

import numpy as np

import cases
import harness
from util import relerr, relerr_elem

pkg = importlib.import_module("lammps-sph-multiphase_b200")


def valid_grid(deck, grid):
    """one ghost layer: every swapped dimension's sub-domain must be at least one ghost cutoff long (comm_brick.cpp:228-230)"""
    for d in range(3):
        if deck.dimension == 2 and d == 2:
            if grid[d] != 1:
                return False
            continue
        prd = deck.boxhi[d] - deck.boxlo[d]
        if (deck.periodicity[d] or grid[d] > 1) and int(deck.cutghost * grid[d] / prd) + 1 > 1:
            return False
    return True


def pick_grid(deck, world, prefer=None):
    """`prefer` if it is valid for this deck, else the LAMMPS grid (ProcMap::onelevel_grid), else any valid factorisation"""
    cands = []
    if prefer is not None and prefer[0] * prefer[1] * prefer[2] == world:
        cands.append(tuple(prefer))
    cands.append(pkg.parallel.proc_grid(world, deck.boxlo, deck.boxhi, deck.dimension))
    cands += pkg.parallel.factor3(world, deck.dimension)
    for g in cands:
        if valid_grid(deck, g):
            return g
    return None


def check_case(name, dist, rank, world, local, grid=None, nsteps=None, balance=None, vs_world=False):
    """-> dict on every rank (filled on rank 0): name, grid, against, ok, err (max over fields, norm-wise), err_elem, detail.
    balance = "x" | "xy" | "xyz": non-uniform bricks from parallel.balance_shift (the `balance 1.05 shift <dims> 10 1.05` command)"""
    api = pkg.load()
    case = cases.CASES[name]
    g = harness.load_golden(name)
    deck = case.deck()
    grid = pick_grid(deck, world, grid)
    if grid is None:
        return dict(name=name, grid=None, against="skipped", ok=True, err=0.0, detail="no valid %d-rank grid for this box" % world)
    nsteps = case.nsteps if nsteps is None else nsteps
    splits = None
    if balance:
        splits = pkg.parallel.balance_shift(g["init_x"], deck.boxlo, deck.boxhi, grid, balance, 10, 1.05)
    brick = pkg.parallel.Brick(world, rank, deck.boxlo, deck.boxhi, deck.dimension, grid, splits)
    nid = pkg.parallel.nccl_id(api, dist)
    sim = pkg.B200Sim(deck, device=local, brick=brick, nccl_id=nid)
    # the reference sequence: run 0 from the initial state, then run N (tests/golden/make_golden.py)
    st = harness.state_from(g, "init_", case.multiphase)
    mine = brick.owns(st["x"])
    sim.set_atoms(**{k: np.ascontiguousarray(v[mine]) for k, v in st.items()})
    sim.setup()
    sim.setup()
    sim.run(nsteps)
    out = sim.get_atoms()
    nl, ng = sim.natoms()
    c = sim.counters()
    sim.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, (out, nl, ng, c["builds"]))
    res = dict(name=name + ("[balance %s]" % balance if balance else ""), grid=list(grid), against="", ok=True, err=0.0, err_elem=0.0, detail="")
    if rank != 0:
        dist.barrier()
        return res
    tags = np.concatenate([o[0]["tag"] for o in gathered])
    order = np.argsort(tags)
    res["atoms_per_rank"] = [o[1] for o in gathered]
    res["ghosts_per_rank"] = [o[2] for o in gathered]
    res["builds"] = [o[3] for o in gathered]
    moving_mp = case.multiphase and "static" not in name
    pc = "phase_change" in str(case.cmds)
    if moving_mp or pc or vs_world or nsteps != case.nsteps:      # fix phase_change decks: every rank walks its own RanPark stream, in the P-rank oracle too
        from pworld import OracleWorld
        w = OracleWorld(case.deck(), world, grid, splits)
        w.set_atoms(**harness.state_from(g, "init_", case.multiphase))
        w.setup(); w.setup(); w.run(nsteps)
        # the ranks' LOCAL ORDER (LAMMPS local indices after migration: exchange()'s hole filling, arrivals appended, comm_brick.cpp:628-664)
        per_rank = [sim.get_atoms(("tag",))["tag"] for sim in w.sims]
        res["local_order_same"] = bool(all(np.array_equal(per_rank[r], gathered[r][0]["tag"]) for r in range(world)))
        want = w.get_atoms(); w.close()
        fields = ["x", "v", "f", "rho", "e", "de", "drho"] + (["colorgradient", "rmass"] if case.multiphase else [])
        res["against"] = "oracle emulating the same %d ranks" % world
        if len(tags) != len(want["tag"]) or not np.array_equal(tags[order], want["tag"]):
            res["ok"] = False; res["detail"] = "particle sets differ: %d atoms vs %d in the P-rank oracle" % (len(tags), len(want["tag"]))
        else:
            if pc:
                res["detail_pc"] = "%d atoms, %d inserted" % (len(tags), len(tags) - len(g["init_tag"]))
            errs = {k: relerr(np.concatenate([o[0][k] for o in gathered])[order], want[k]) for k in fields}
            elem = {k: relerr_elem(np.concatenate([o[0][k] for o in gathered])[order], want[k]) for k in fields}
            res["err"] = float(max(errs.values())); res["err_elem"] = float(max(elem.values()))
            res["ok"] = bool(res["err"] <= 100 * case.tol_traj)
            res["detail"] = {k: "%.1e" % v for k, v in errs.items()}
    else:
        ref_order = np.argsort(g["sN_tag"])
        res["against"] = "1-rank reference fixture"
        if len(tags) != len(g["sN_tag"]) or not np.array_equal(tags[order], g["sN_tag"][ref_order]):
            res["ok"] = False; res["detail"] = "particle sets differ"
        else:
            fields = ["x", "v", "f", "rho", "e", "de", "drho"] + (["colorgradient", "rmass"] if case.multiphase else [])
            errs = {k: relerr(np.concatenate([o[0][k] for o in gathered])[order], g["sN_" + k][ref_order]) for k in fields}
            elem = {k: relerr_elem(np.concatenate([o[0][k] for o in gathered])[order], g["sN_" + k][ref_order]) for k in fields}
            res["err"] = float(max(errs.values())); res["err_elem"] = float(max(elem.values()))
            res["ok"] = bool(res["err"] <= 10 * case.tol_traj)
            res["detail"] = {k: "%.1e" % v for k, v in errs.items()}
    dist.barrier()
    return res


def format_result(r):
    order = "" if "local_order_same" not in r else ("  local order of every rank as in the oracle" if r["local_order_same"] else "  LOCAL ORDER DIFFERS")
    return "%-24s grid %s vs %s: %s  err %.1e (element-wise %.1e)  atoms/rank %s ghosts %s builds %s%s  %s" % (
        r["name"], r["grid"], r["against"], "OK" if r["ok"] else "FAIL", r.get("err", 0.0), r.get("err_elem", 0.0),
        r.get("atoms_per_rank"), r.get("ghosts_per_rank"), r.get("builds"), order, r.get("detail"))
