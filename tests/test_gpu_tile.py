"""GPU (B200): the shared-memory tile path (b200_tile.cuh) against the row path (b200_pair.cuh) of the
same library on the same inputs.  The row path gathers every neighbor record from global memory and
walks 32-bit rows; the tile path stages records in shared memory and walks 16-bit slot rows built with
an fp32 pre-decision + exact fallback.  Both must give the same neighbor lists (bit-exact) and the same
per-atom fields up to summation order and the 1e-13 sqrt / reciprocal of the uniform force body (2e-11 over the runs), including periodic decks (ghost candidates) and decks
whose coefficients depend on the type pair (non-uniform tables)."""
import importlib
import os

import numpy as np
import pytest

import cases
import harness

pkg = importlib.import_module("lammps-sph-multiphase_b200")
pytestmark = pytest.mark.gpu


def _run(name, nsteps, env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        case = cases.CASES[name]
        g = harness.load_golden(name)
        sim = pkg.B200Sim(case.deck())
        sim.set_atoms(**harness.state_from(g, "init_", case.multiphase))
        sim.setup()
        sim.run(nsteps)
        out = (sim.get_atoms(), sim.neighbor_list(), sim.counters())
        sim.close()
        return out
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("name,nsteps", [("dam3d", 12), ("dam2d", 20), ("dam2d_morris", 20), ("heat3d", 25), ("heat2d_rhosum", 20), ("gas3d", 10)])
@pytest.mark.parametrize("variant", [{}, {"B200_TILE_NOUNI": "1"}, {"B200_TILE_SPLIT": "1"}, {"B200_TILE_SPLIT": "4"}])
def test_tile_path_equals_row_path(name, nsteps, variant):
    a, na, ca = _run(name, nsteps, dict(variant))
    b, nb, cb = _run(name, nsteps, {"B200_NO_TILE": "1"})
    assert ca["builds"] == cb["builds"]
    for p, q in zip(na, nb):
        assert np.array_equal(p, q), "neighbor lists differ"
    for k in ("x", "v", "vest", "f", "rho", "drho", "e", "de"):       # the uniform force body uses 1e-13-accurate sqrt / reciprocal (b200_tile.cuh)
        assert harness.relerr(a[k], b[k]) < 2e-11, (name, k, harness.relerr(a[k], b[k]))


@pytest.mark.parametrize("name,nsteps", [("droplet3d", 10), ("droplet2d", 20), ("droplet3d_heat", 8), ("droplet2d_pcheat_skin", 20), ("bubble3d", 8),
                                         ("bubble2d_thermostat", 12), ("kat_surfacetension", 1)])
@pytest.mark.parametrize("variant", [{}, {"B200_TILE_NOUNI": "1"}])
def test_multiphase_tile_path_equals_row_path(name, nsteps, variant):
    """multiphase styles (ownership flags, ghost-row tiles, reverse halo, fix phase_change reading tile rows)"""
    a, na, ca = _run(name, nsteps, dict(variant))
    b, nb, cb = _run(name, nsteps, {"B200_NO_TILE_MP": "1"})
    assert ca["builds"] == cb["builds"] and ca["inserted"] == cb["inserted"]
    assert ca["launches"] != cb["launches"]
    for p, q in zip(na, nb):
        assert np.array_equal(p, q), "neighbor lists differ"
    for k in ("x", "v", "vest", "f", "rho", "drho", "e", "de", "colorgradient", "rmass"):
        assert harness.relerr(a[k], b[k]) < 1e-10, (name, k, harness.relerr(a[k], b[k]))


@pytest.mark.parametrize("name,nsteps,cap", [("dam3d", 12, 900), ("dam3d", 12, 64), ("droplet3d", 10, 400), ("droplet3d", 10, 64), ("heat2d_rhosum", 20, 40)])
def test_small_tiles_and_fallback(name, nsteps, cap):
    """a shared-memory budget that allows only one or two cells per tile, or (cap 64 / 40) not even one cell:
    the plan must cut smaller tiles, or hand the deck to the row path, without changing the results"""
    a, na, ca = _run(name, nsteps, {"B200_TILE_SLOTCAP": str(cap)})
    b, nb, cb = _run(name, nsteps, {})
    assert ca["builds"] == cb["builds"]
    for p, q in zip(na, nb):
        assert np.array_equal(p, q), "neighbor lists differ"
    for k in ("x", "v", "vest", "f", "rho", "drho", "e", "de"):
        assert harness.relerr(a[k], b[k]) < 1e-10, (name, k, harness.relerr(a[k], b[k]))


@pytest.mark.parametrize("name,nsteps", [("heat3d", 25), ("heat2d_rhosum", 20), ("gas3d", 10)])
def test_halo_overlap_equals_plain_order(name, nsteps):
    """B200_OVERLAP=1: interior tiles run while the (here: periodic self-) halo is in flight on a second stream"""
    a, na, ca = _run(name, nsteps, {"B200_OVERLAP": "1"})
    b, nb, cb = _run(name, nsteps, {})
    assert ca["launches"] > cb["launches"]
    for p, q in zip(na, nb):
        assert np.array_equal(p, q)
    for k in ("x", "v", "vest", "f", "rho", "drho", "e", "de"):
        assert harness.relerr(a[k], b[k]) < 1e-12, (name, k, harness.relerr(a[k], b[k]))


def test_tile_path_is_the_one_that_runs():
    """single-phase decks must take the tile kernels (the launch counter differs from the row path's)"""
    a, _, ca = _run("dam3d", 5, {})
    b, _, cb = _run("dam3d", 5, {"B200_NO_TILE": "1"})
    assert ca["launches"] != cb["launches"]


def _run_c2(scale, nsteps, env):
    import bench
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        atoms, params = bench.dam_break_3d(scale)
        sim = pkg.B200Sim(bench.make_deck(pkg, params))
        sim.set_atoms(**atoms)
        sim.setup(); sim.run(nsteps)
        out = sim.get_atoms()
        sim.close()
        return out
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("name,nsteps", [("dam3d", 12), ("dam2d", 40), ("gas3d", 15), ("gas3d_shrink", 45), ("c2@0.35", 14)])
def test_per_tile_zone_flags_equal_the_global_flag(name, nsteps):
    """Mid / far rows hold only pairs beyond the cutoff until 2 * dmax reaches their margin, and such pairs add exactly 0.0:
    switching them on per tile (local displacement bound, the default) or for all tiles at once (B200_ZONE_GLOBAL=1) must give
    bitwise identical fields."""
    if name.startswith("c2@"):
        a = _run_c2(float(name[3:]), nsteps, {})
        b = _run_c2(float(name[3:]), nsteps, {"B200_ZONE_GLOBAL": "1"})
    else:
        a = _run(name, nsteps, {})[0]
        b = _run(name, nsteps, {"B200_ZONE_GLOBAL": "1"})[0]
    for k in ("x", "v", "vest", "f", "rho", "drho", "e", "de"):
        assert np.array_equal(a[k], b[k]), (name, k, harness.relerr(a[k], b[k]))


@pytest.mark.parametrize("name", ["cavity2d_rhosum", "cavity2d"])
def test_stale_setup_returns_to_the_tile_path(name):
    """non-zero initial velocities in a periodic single-phase box: the ghosts carry a stale vest through the setup force evaluation, which
    is taken on the row path (reference pair orientation, reverse halo); the run itself is back on the tile kernels and equals a run
    that stays on the row path"""
    a, na, ca = _run(name, 30, {})
    b, nb, cb = _run(name, 30, {"B200_STALE_SETUP_STAYS_ON_ROWS": "1"})
    c, nc, cc = _run(name, 30, {"B200_NO_TILE": "1"})
    assert ca["builds"] == cb["builds"] == cc["builds"]
    assert ca["launches"] != cb["launches"]          # tile kernels against row kernels over the 30 steps
    for p, q, r in zip(na, nb, nc):
        assert np.array_equal(p, q) and np.array_equal(p, r), "neighbor lists differ"
    for k in ("x", "v", "vest", "f", "rho", "drho", "e", "de"):
        assert harness.relerr(a[k], b[k]) < 2e-11, (name, k, harness.relerr(a[k], b[k]))
        assert harness.relerr(b[k], c[k]) == 0.0, (name, k)
