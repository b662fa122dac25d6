"""CPU: the deck texts of tests/test_gpu_lammps_shell.py (the repo's own shipped-style decks: dam break, heat conduction, droplet, bubble
with phase change, shock tubes, sph/lj, variable body forces, fix print / fix ave/spatial, a refused fix, two runs, the water_collapse
thermo line) through `lmp_b200`'s shells with the oracle standing in for the CUDA library (tests/shipped.py) against `lmp_serial` --
the same test functions, with the library behind the C-ABI swapped by LD_PRELOAD.  Whatever the shells do wrong on the host shows here
before a GPU is involved."""
import os
import subprocess

import pytest

import shipped
import test_gpu_lammps_shell as G

pytestmark = pytest.mark.skipif(not shipped.available(), reason="needs /root/reference, oracle/_ref/lmp_serial and lmp_b200")


@pytest.fixture
def stand_in(monkeypatch):
    shim = shipped.build_shim()
    run0, sprun0 = G.run, subprocess.run

    def run(exe, args, workdir, text):
        if exe == G.B200:
            monkeypatch.setenv("LD_PRELOAD", shim)
        try:
            return run0(exe, args, workdir, text)
        finally:
            monkeypatch.delenv("LD_PRELOAD", raising=False)

    def sprun(cmd, **kw):          # the refusal test calls subprocess.run itself
        if cmd and cmd[0] == G.B200 and "env" not in kw and "LD_PRELOAD" not in os.environ:
            kw["env"] = dict(os.environ, LD_PRELOAD=shim)
        return sprun0(cmd, **kw)

    monkeypatch.setattr(G, "run", run)
    monkeypatch.setattr(G.subprocess, "run", sprun)


PARAMS = [m.args for m in G.test_same_deck_reference_vs_b200.pytestmark if m.name == "parametrize"][0][1]


@pytest.mark.parametrize("name,nsteps,tol", PARAMS, ids=[p[0] for p in PARAMS])
def test_same_deck_through_the_shells(name, nsteps, tol, tmp_path, stand_in):
    G.test_same_deck_reference_vs_b200.__wrapped__(name, nsteps, tol, tmp_path) if hasattr(G.test_same_deck_reference_vs_b200, "__wrapped__") \
        else G.test_same_deck_reference_vs_b200(name, nsteps, tol, tmp_path)


@pytest.mark.parametrize("fn", ["test_host_end_of_step_fixes_fire", "test_poiseuille_profile_by_fix_ave_spatial", "test_unsupported_stepping_fix_is_refused",
                                "test_phase_change_state_survives_a_second_run", "test_water_collapse_thermo_keywords"])
def test_shell_behaviour(fn, tmp_path, stand_in):
    getattr(G, fn)(tmp_path)
