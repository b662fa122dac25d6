"""GPU: the drop-in itself.  The SAME deck text is run by the unmodified reference binary
(oracle/_ref/lmp_serial) and by the reference + USER-B200 shells (lammps/_build/lmp_b200 -sf b200),
and the final per-atom dumps (17 significant digits, sorted by id) are compared.
Both binaries are built where /root/reference exists and travel to the GPU box."""
import os
import re
import subprocess

import numpy as np
import pytest

import cases
from util import relerr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "lmp_serial")
B200 = os.path.join(ROOT, "lammps-sph-multiphase_b200", "lammps", "_build", "lmp_b200")
pytestmark = pytest.mark.gpu

COLS = "id type x y z vx vy vz fx fy fz c_crho c_ce"
FMT = "%d %d " + " ".join(["%.17g"] * 11)


def deck_text(case, nsteps):
    return "\n".join([re.sub(r"atom_modify map array( sort \S+ \S+)?", "atom_modify map array", case.header_text()), case.create, case.lammps_text(),
                      "compute crho all meso_rho/atom", "compute ce all meso_e/atom", "thermo 10",
                      "thermo_style custom step press pxx pyy pxy", "thermo_modify format float %.15g norm no",
                      "dump dfin all custom %d dump.final %s" % (nsteps, COLS), 'dump_modify dfin sort id format "%s"' % FMT,
                      "run %d" % nsteps, ""])


def run(exe, args, workdir, text):
    os.makedirs(workdir, exist_ok=True)
    with open(os.path.join(workdir, "deck.lmp"), "w") as f:
        f.write(text)
    p = subprocess.run([exe] + args + ["-in", "deck.lmp", "-log", "log.lammps"], cwd=workdir, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "ERROR" not in p.stdout, p.stdout[-3000:] + p.stderr[-2000:]
    rows = []
    with open(os.path.join(workdir, "dump.final")) as f:
        lines = f.read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith("ITEM: ATOMS")][-1]      # last snapshot = end of the run
    for l in lines[start + 1:]:
        rows.append([float(v) for v in l.split()])
    return np.array(rows), p.stdout


def thermo_rows(stdout):
    """the numeric rows under the `Step Press Pxx Pyy Pxy` header"""
    rows, on = [], False
    for l in stdout.splitlines():
        t = l.split()
        if t[:2] == ["Step", "Press"]:
            on = True; continue
        if on:
            try:
                rows.append([float(v) for v in t])
                assert len(t) == 5
            except (ValueError, AssertionError):
                on = False
    return np.array(rows)


@pytest.mark.parametrize("name,nsteps,tol", [("dam2d", 40, 1e-9), ("heat2d", 60, 1e-10), ("droplet3d", 10, 1e-9), ("bubble2d", 20, 1e-8), ("shock2d", 30, 1e-9), ("heat2d_setmesode", 30, 1e-10), ("dam2d_dtreset", 30, 1e-9),
                                            ("shock2d_shrink", 40, 1e-9), ("shock3d_shrink", 25, 1e-9), ("lj2d", 40, 1e-9),
                                            ("dam2d_addforce", 30, 1e-9), ("heat2d_setmeso_var", 30, 1e-10), ("poiseuille2d", 50, 1e-9)])
def test_same_deck_reference_vs_b200(name, nsteps, tol, tmp_path):
    if not (os.path.exists(REF) and os.path.exists(B200)):
        pytest.skip("lmp_serial / lmp_b200 not built (they are built only where /root/reference exists)")
    case = cases.CASES[name]
    text = deck_text(case, nsteps)
    a, a_out = run(REF, [], str(tmp_path / "ref"), text)
    b, out = run(B200, ["-sf", "b200"], str(tmp_path / "b200"), text)
    assert "B200 engine" in out
    assert a.shape == b.shape, "particle counts differ: %s vs %s" % (a.shape, b.shape)
    assert np.array_equal(a[:, :2], b[:, :2])
    names = COLS.split()
    errs = {}
    for lo, hi, nm in ((2, 5, "x"), (5, 8, "v"), (8, 11, "f"), (11, 12, "rho"), (12, 13, "e")):
        errs[nm] = relerr(b[:, lo:hi], a[:, lo:hi])
    assert all(v <= tol for v in errs.values()), errs
    # thermo pressure (needs the pair virial of the engine on thermo steps): every printed line must agree
    pa, pb = thermo_rows(a_out), thermo_rows(out)
    assert pa.shape == pb.shape and len(pa) >= 2, (pa.shape, pb.shape)
    assert np.array_equal(pa[:, 0], pb[:, 0])
    scale = np.abs(pa[:, 1:]).max()
    assert np.abs(pa[:, 1:] - pb[:, 1:]).max() <= 100 * tol * max(scale, 1e-300), (pa[-1], pb[-1])


def _both(tmp_path, text):
    if not (os.path.exists(REF) and os.path.exists(B200)):
        pytest.skip("lmp_serial / lmp_b200 not built (they are built only where /root/reference exists)")
    a, a_out = run(REF, [], str(tmp_path / "ref"), text)
    b, b_out = run(B200, ["-sf", "b200"], str(tmp_path / "b200"), text)
    assert "B200 engine" in b_out
    return a, a_out, b, b_out


def test_host_end_of_step_fixes_fire(tmp_path):
    """fix print (END_OF_STEP, no /b200 variant) must fire on its steps with the values of THAT step: VerletB200::run ends a
    device-resident segment on every `nevery`, refreshes the host arrays and calls modify->end_of_step() (verlet.cpp:300)"""
    case = cases.CASES["dam2d"]
    nsteps = 30
    extra = "\n".join(["variable s equal step", "variable k equal vcm(all,x)", "variable xc equal xcm(all,x)",
                       'fix pr all print 7 "PRINTED ${s} ${k} ${xc}"'])
    text = deck_text(case, nsteps).replace("run %d" % nsteps, extra + "\nrun %d" % nsteps)
    a, a_out, b, b_out = _both(tmp_path, text)
    pa = [l.split()[1:] for l in a_out.splitlines() if l.startswith("PRINTED")]
    pb = [l.split()[1:] for l in b_out.splitlines() if l.startswith("PRINTED")]
    assert len(pa) == len(pb) and len(pa) >= 4, (pa, pb)
    for ra, rb in zip(pa, pb):
        assert ra[0] == rb[0]
        assert relerr(np.array([float(v) for v in rb[1:]]), np.array([float(v) for v in ra[1:]])) <= 1e-5     # fix print writes %g-style 6 digits... compare at that precision
    assert relerr(b[:, 2:5], a[:, 2:5]) <= 1e-9


def test_poiseuille_profile_by_fix_ave_spatial(tmp_path):
    """the second half of examples/USER/sph/poiseuille/poiseuille.lmp: the velocity profile is collected by fix ave/spatial (a host
    END_OF_STEP fix fed by a compute and an atom-style variable) while fix addforce/b200 drives the flow on the device"""
    case = cases.CASES["poiseuille2d"]
    nsteps = 40
    extra = "\n".join(["compute vxav all reduce ave vx", "variable vx_cm atom vx-c_vxav",
                       "fix av_vx all ave/spatial 5 4 20 y center 0.05 v_vx_cm file vx.av ave one units reduced"])
    text = deck_text(case, nsteps).replace("run %d" % nsteps, extra + "\nrun %d" % nsteps)
    a, a_out, b, b_out = _both(tmp_path, text)
    prof = []
    for d in ("ref", "b200"):
        rows = [[float(v) for v in l.split()] for l in open(str(tmp_path / d / "vx.av")) if not l.startswith("#")]
        prof.append(np.array([r for r in rows if len(r) == 4]))          # chunk rows: index coord ncount value
    assert prof[0].shape == prof[1].shape and len(prof[0]) >= 20, (prof[0].shape, prof[1].shape)
    assert np.array_equal(prof[0][:, :3], prof[1][:, :3])
    assert relerr(prof[1][:, 3], prof[0][:, 3]) <= 1e-5                   # the file holds 6 significant digits
    assert relerr(b[:, 5:8], a[:, 5:8]) <= 1e-9


def test_unsupported_stepping_fix_is_refused(tmp_path):
    """a fix with per-step hooks and no /b200 variant is an error, not a silent no-op"""
    if not os.path.exists(B200):
        pytest.skip("lmp_b200 not built")
    case = cases.CASES["dam2d"]
    text = deck_text(case, 5).replace("run 5", "fix bad all momentum 1 linear 1 1 1\nrun 5")
    os.makedirs(str(tmp_path / "b"), exist_ok=True)
    with open(str(tmp_path / "b" / "deck.lmp"), "w") as f:
        f.write(text)
    p = subprocess.run([B200, "-sf", "b200", "-in", "deck.lmp", "-log", "none"], cwd=str(tmp_path / "b"), capture_output=True, text=True, timeout=300)
    assert "has no /b200 variant" in p.stdout + p.stderr


def test_phase_change_state_survives_a_second_run(tmp_path):
    """`run 10` twice == what the reference does with the same two commands: fix phase_change keeps next_reneighbor and its
    RNG stream across runs (fix_phase_change.cpp:116,345), so must the engine behind VerletB200::configure"""
    case = cases.CASES["bubble2d"]
    text = deck_text(case, 20).replace("run 20", "run 10\nrun 10")
    a, a_out, b, b_out = _both(tmp_path, text)
    assert a.shape == b.shape, "particle counts differ: %s vs %s" % (a.shape, b.shape)
    assert np.array_equal(a[:, :2], b[:, :2])
    for lo, hi in ((2, 5), (5, 8), (8, 11), (11, 12), (12, 13)):
        assert relerr(b[:, lo:hi], a[:, lo:hi]) <= 1e-8


def test_water_collapse_thermo_keywords(tmp_path):
    """examples/USER/sph/water_collapse/water_collapse.lmp:42 prints `ke c_esph v_etot f_gfix press time f_dtfix`: the potential energy of
    fix gravity (FixGravity::compute_scalar), the elapsed time under the variable timestep (Update::atime, thermo.cpp:1500) and the last
    step fix dt/reset changed it on (FixDtReset::compute_scalar) -- with the force, the timestep and the time bookkeeping on the device"""
    case = cases.CASES["dam2d_dtreset"]
    text = "\n".join([case.header_text(), case.create, case.lammps_text(),
                      "compute ce all meso_e/atom", "compute esph all reduce sum c_ce", "compute cke all ke", "variable etot equal c_esph+c_cke+f_f1",
                      "thermo 7", "thermo_style custom step ke c_esph v_etot f_f1 press time f_f4", "thermo_modify format float %.15g norm no",
                      "dump dfin all custom 35 dump.final %s" % COLS.replace("c_crho c_ce", "vx vy"), 'dump_modify dfin sort id format "%s"' % FMT,
                      "run 35", "run 14", ""])
    a, a_out, b, b_out = _both(tmp_path, text)

    def rows(out):
        r, on = [], False
        for l in out.splitlines():
            t = l.split()
            if t[:2] == ["Step", "KinEng"]:
                on = True; continue
            if on:
                try:
                    r.append([float(v) for v in t]); assert len(t) == 8
                except (ValueError, AssertionError):
                    on = False
        return np.array(r)
    ra, rb = rows(a_out), rows(b_out)
    assert ra.shape == rb.shape and len(ra) >= 9, (ra.shape, rb.shape)
    assert np.array_equal(ra[:, 0], rb[:, 0]) and np.array_equal(ra[:, 7], rb[:, 7]), (ra[:, 7], rb[:, 7])      # steps, f_dtfix
    assert len(set(ra[:, 7])) > 2                                                                              # the timestep did change
    for c, nm in ((1, "ke"), (2, "esph"), (3, "etot"), (4, "f_gfix"), (5, "press"), (6, "time")):
        s = np.abs(ra[:, c]).max()
        assert np.abs(ra[:, c] - rb[:, c]).max() <= 1e-8 * max(s, 1e-300), (nm, ra[:, c], rb[:, c])
