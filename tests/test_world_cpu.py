"""CPU: the P-rank emulation of the oracle (tests/pworld.py) against the 1-rank reference fixtures.
Decks whose result does not depend on the decomposition (single-phase decks: fresh ghost data; static multiphase decks) must
come out as the reference computed them on one rank, for several brick grids -- this pins the emulated exchange / borders /
forward / reverse collectives.  Moving multiphase decks only have to stay close (stale ghost rho): their P-rank
values are what the multi-GPU engine is compared with (tests/mgpu_check.py)."""
import numpy as np
import pytest

import cases
import harness
from pworld import OracleWorld
from util import relerr


def _run(name, world, grid=None):
    case = cases.CASES[name]
    g = harness.load_golden(name)
    w = OracleWorld(case.deck(), world, grid)
    w.set_atoms(**harness.state_from(g, "init_", case.multiphase))
    w.setup(); w.setup()              # the reference sequence: run 0, then run N (tests/golden/make_golden.py)
    w.run(case.nsteps)
    out, nat, builds = w.get_atoms(), w.natoms(), w.builds()
    w.close()
    return case, g, out, nat, builds


@pytest.mark.parametrize("name,world,grid", [("dam3d", 2, None), ("dam3d", 4, None), ("dam3d", 3, (3, 1, 1)), ("dam2d", 2, None), ("dam2d", 4, (2, 2, 1)),
                                             ("heat3d", 2, None), ("heat2d_rhosum", 2, None), ("heat2d_rhosum", 4, (2, 2, 1)), ("gas3d", 2, None), ("gas3d", 4, None),
                                             ("droplet3d_static", 2, None), ("droplet2d_static", 2, None), ("droplet2d_static", 4, (2, 2, 1))])
def test_world_reproduces_one_rank_fixture(name, world, grid):
    # not in the list: the two-type shock decks.  PairSPHIdealGas leaves viscosity[j][i] unset, so a cross-type pair
    # depends on which atom the half list puts first, i.e. on the local index order -- and CommBrick::exchange fills the hole of a
    # migrated atom with the rank's last atom, so the reference itself gives different forces on 1 and on 2 ranks there
    # (measured here: 2e-2 in f around the migration sites of shock3d at 2 ranks).
    case, g, out, nat, builds = _run(name, world, grid)
    ref = np.argsort(g["sN_tag"])
    assert np.array_equal(out["tag"], g["sN_tag"][ref]), "particles lost or duplicated: %s" % (nat,)
    assert len(set(builds)) == 1 and builds[0] == int(g["sN_nbuilds"]), (builds, int(g["sN_nbuilds"]))
    fields = ["x", "v", "f", "rho", "e", "de", "drho"] + (["colorgradient", "rmass"] if case.multiphase else [])
    errs = {k: relerr(out[k], g["sN_" + k][ref]) for k in fields}
    assert all(v <= 10 * case.tol_traj for v in errs.values()), errs


@pytest.mark.parametrize("name", ["droplet3d", "droplet2d"])
def test_world_moving_multiphase_stays_close(name):
    case, g, out, nat, builds = _run(name, 2)
    ref = np.argsort(g["sN_tag"])
    assert np.array_equal(out["tag"], g["sN_tag"][ref])
    assert relerr(out["x"], g["sN_x"][ref]) < 3e-2 and relerr(out["rho"], g["sN_rho"][ref]) < 3e-2
