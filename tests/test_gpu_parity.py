"""GPU (B200): the CUDA engine, driven through the C-ABI, against
 (a) the fixtures dumped from the real reference build (tests/golden, all cases), and
 (b) the CPU oracle run live on the same inputs.
Tolerances (north star): neighbor lists and particle counts bit-exact; rho, forces, de/dt
within 1e-10 relative (max|a-b|/max|b|) for one force evaluation; trajectories looser as
round-off differences grow (the reference's own sums are order dependent)."""
import importlib

import numpy as np
import pytest

import cases
import harness

pkg = importlib.import_module("lammps-sph-multiphase_b200")
pytestmark = pytest.mark.gpu

TOL_STEP = 1e-10
ALL = [n for n, c in cases.CASES.items() if c.engine]


@pytest.mark.parametrize("name", ALL)
def test_engine_matches_reference_fixture(name):
    e0, eN = harness.run_case(pkg.B200Sim, name, tol_step=TOL_STEP)
    print(name, "run0", {k: "%.1e" % v for k, v in e0.items()}, "runN", {k: "%.1e" % v for k, v in eN.items()})


def test_engine_refuses_what_it_does_not_implement():
    """no silent fallback: sph/lj without a full-list sub-style (LAMMPS would build half_bin_newton lists in another order, and the
    style's result depends on the list order) fails at setup with a message"""
    case = cases.CASES["lj3d"]
    g = harness.load_golden("lj3d")
    deck = case.deck()
    deck.styles = [s for s in deck.styles if s.name != "sph/rhosum"]
    sim = pkg.B200Sim(deck)
    sim.set_atoms(**harness.state_from(g, "init_", False))
    with pytest.raises(RuntimeError) as e:
        sim.setup()
    assert "full-list sub-style" in str(e.value)
    sim.close()


@pytest.mark.parametrize("name", ["dam3d", "droplet3d", "heat2d", "bubble3d"])
def test_engine_matches_oracle_live(name):
    case = cases.CASES[name]
    g = harness.load_golden(name)
    sims = [mk(case.deck()) for mk in (pkg.B200Sim, harness.oracle_sim)]
    for s in sims:
        s.set_atoms(**harness.state_from(g, "init_", case.multiphase))
        s.setup()
    for chunk in range(3):
        outs = []
        for s in sims:
            s.run(4)
            outs.append((s.get_atoms(), s.neighbor_list(), s.natoms()))
        (a, na, ca), (b, nb, cb) = outs
        assert ca == cb
        for p, q in zip(na, nb):
            assert np.array_equal(p, q)
        for k in ("x", "v", "vest", "f", "rho", "drho", "e", "de"):
            assert harness.relerr(a[k], b[k]) < 1e-9, (name, chunk, k, harness.relerr(a[k], b[k]))
    for s in sims:
        s.close()


def test_elapsed_time_under_fix_dt_reset_matches_oracle():
    """Update::atime / atimestep and FixDtReset::laststep (thermo `time`, f_ID of the fix) advance on the device exactly as
    FixDtReset::end_of_step + Update::update_time advance them (fix_dt_reset.cpp:175-181, update.cpp:480-484); the oracle's
    restatement is itself checked against lmp_serial's thermo output of the shipped water_collapse deck (tests/test_shell_shipped_cpu.py)"""
    case = cases.CASES["dam2d_dtreset"]
    g = harness.load_golden("dam2d_dtreset")
    sims = [mk(case.deck()) for mk in (pkg.B200Sim, harness.oracle_sim)]
    for s in sims:
        s.set_atoms(**harness.state_from(g, "init_", case.multiphase))
        s.set_time(0.125, 0, 0)
        s.setup()
    seen = set()
    for chunk in range(4):
        for s in sims:
            s.run(7)
        (ta, sa, la), (tb, sb, lb) = sims[0].time(), sims[1].time()
        assert (sa, la) == (sb, lb), (chunk, sa, la, sb, lb)
        # the timestep follows the forces (1e-12 per step between engine and oracle), the elapsed time sums timesteps
        assert abs(ta - tb) <= 1e-10 * (tb - 0.125), (chunk, ta, tb)
        assert abs(sims[0].timestep() - sims[1].timestep()) <= 1e-10 * sims[1].timestep()
        seen.add(la)
    assert len(seen) > 1 and tb > 0.125          # the timestep did change along the way
    for s in sims:
        s.close()
