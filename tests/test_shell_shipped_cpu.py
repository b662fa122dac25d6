"""CPU: the reference's shipped example decks through `lmp_b200 -sf b200` with the oracle standing in for the CUDA
library behind the C-ABI (tests/shipped.py), against the unmodified `lmp_serial` on the same text."""
import os

import pytest

import shipped
from shipped import Shipped

pytestmark = pytest.mark.skipif(not shipped.available(), reason="needs /root/reference, oracle/_ref/lmp_serial and lmp_b200 (built by __graft_entry__.build() where the reference exists)")

D = ["-var", "dname", "data"]
# slub/infslab.lmp includes a par.lmp its README generates with maxima (absent here): the same parameter list (infslab.mac `env`), written by hand
PAR = "printf 'variable xm equal 0.5\nvariable Lx equal 1.0\nvariable cv_r equal 4.179\nvariable cv_l equal 1.0\nvariable rho_r equal 1000\n" \
      "variable rho_l equal 1.226\nvariable k_l equal 0.0254\nvariable k_r equal 5\nvariable tau equal 0.5\nvariable t_r equal 2\nvariable t_l equal 1\n' > par.lmp"

# every shipped deck that runs without maxima-generated input (multiphase_two_atoms/*.lmp need it; their closed forms are the kat_* fixtures).
# Sizes: the -var arguments of each run.sh with nx reduced; `cap` bounds every `run`.
CASES = [
    # atom-style variable body force (fix addforce v_bodyfx), fix ave/spatial profile, two runs -- verbatim, full length (150 atoms)
    Shipped("poiseuille", "poiseuille", "poiseuille.lmp", var=D, cap=900, files=["data/vx.av"]),
    # fix addforce, fix setforce, two fix ave/spatial, write_data
    Shipped("flow_around_cylinder", "flow_around_cylinder", "flow.lmp", var=D, cap=60, pre=["mkdir -p data"]),
    # fix phase_change + variable setmeso (atom-style) + unfix + count()/xcm()/bound() variables in fix print 1 + `run N pre no post no every M "if ..."`
    Shipped("bubble_on_wall", "bubble_on_wall", "bubble.lmp", var=D, cap=60, subs=[(r"every 1000", "every 20")], files=["data/rg.dat"]),
    # three runs with `velocity all set` and a changed pair_coeff in between, setmeso noregion, fix phase_change in the last
    Shipped("bubble_random", "bubble_random", "bubble.lmp", var=["-var", "nx", "12", "-var", "ndim", "3"] + D, cap=30),
    Shipped("bubble_growth", "bubble_growth", "bubble.lmp", var=["-var", "nx", "12", "-var", "ndim", "3"] + D, cap=30),
    # three-phase wetting decks: fix setforce on the wall, compute gyration / reduce through fix print
    Shipped("contact_angle", "contact_angle", "droplet.lmp", var=["-var", "icase", "2", "-var", "nx", "41"] + D, cap=40, pre=["mkdir -p data"], files=["data/com.dat"]),
    Shipped("droplet_grid", "droplet_grid", "droplet.lmp", var=["-var", "icase", "2", "-var", "nx", "42"] + D, cap=40, pre=["mkdir -p data"],
            files=["data/rg.dat", "data/cm.dat"]),
    Shipped("square_to_sphere", "square_to_sphere", "droplet.lmp", var=["-var", "ndim", "3", "-var", "nx", "14"] + D, cap=30, files=["data/rg.dat"]),
    # the decks' other documented cases: icase 1 (no wall), the 2-D square (cylinder.lmp instead of cube.lmp), the 2-D bubble
    Shipped("contact_angle_case1", "contact_angle", "droplet.lmp", var=["-var", "icase", "1", "-var", "nx", "41"] + D, cap=30, pre=["mkdir -p data"], files=["data/com.dat"]),
    Shipped("droplet_grid_case1", "droplet_grid", "droplet.lmp", var=["-var", "icase", "1", "-var", "nx", "42"] + D, cap=30, pre=["mkdir -p data"],
            files=["data/rg.dat", "data/cm.dat"]),
    Shipped("square_to_sphere_2d", "square_to_sphere", "droplet.lmp", var=["-var", "ndim", "2", "-var", "nx", "40"] + D, cap=40, files=["data/rg.dat"]),
    Shipped("bubble_random_2d", "bubble_random", "bubble.lmp", var=["-var", "nx", "40", "-var", "ndim", "2"] + D, cap=40),
    # sph/taitwater/morris alone (half_bin_newton lists), `pair_coeff 2 3 none`, a driver strip that starts with a velocity in a periodic box
    Shipped("cavity_flow", "cavity_flow", "cavity_flow.lmp", cap=400),
    # read_data, fix gravity + its potential energy f_gfix, fix dt/reset + f_dtfix and the thermo keyword `time`, enforce2d, press
    Shipped("water_collapse", "water_collapse", "water_collapse.lmp", cap=400),
    Shipped("heat2d", "heatconduction", "sph_heat_conduction_2d.lmp", cap=160),
    Shipped("heat3d", "heatconduction", "sph_heat_conduction_3d.lmp", cap=40),
    # boundary s p p (shrink-wrapped box), sph/idealgas
    Shipped("shock2d", "shock_tube", "shock2d.lmp", cap=100),
    Shipped("shock3d", "shock_tube", "shock3d.lmp", cap=20),
    Shipped("slub", "slub", "infslab.lmp", cap=60, pre=[PAR]),
]


# multiphase_two_atoms/*.lmp include an in.atoms (and in.vars) that `maxima -b <deck>.mac` prints from the lists at the top of each .mac
# (x, type; gamma, soundspeed, eta, rbackground, rho0): written here from those same lists.  The decks print per-atom results through fix print.
def _atoms(xs, types):
    return "printf '%s' > in.atoms" % "".join("create_atoms %d single %s units box\n" % (t, " ".join("%g" % v for v in x)) for x, t in zip(xs, types))


T3 = [[5, 5, 5], [5.5, 5, 5], [5, 5, 4.8]]
CASES += [
    Shipped("two_atoms_colorgradient", "multiphase_two_atoms", "colorgradient.lmp", cap=1, pre=[_atoms(T3, [1, 2, 2])], files=["outpt.dat"]),
    Shipped("two_atoms_rhosum", "multiphase_two_atoms", "sph_rhosum_multiphase.lmp", cap=1, pre=[_atoms(T3, [1, 2, 2])], files=["outpt.dat"]),
    Shipped("two_atoms_taitwater", "multiphase_two_atoms", "sph_taitwater_multiphase.lmp", cap=1, files=["outpt.dat"],
            pre=[_atoms(T3, [1, 2, 2]), "printf 'variable gamma equal 1\nvariable soundspeed equal 1\nvariable eta equal 0\nvariable rbackground equal 0.5\nvariable rho0 equal 1\n' > in.vars"]),
    Shipped("two_atoms_surfacetension", "multiphase_two_atoms", "surfacetension.lmp", cap=1, files=["output.dat"],
            pre=[_atoms([[4.6, 5.3, 5], [5.5, 5, 5.2], [5.0, 5.0, 5.0]], [1, 2, 2])]),
    Shipped("two_atoms_heat_phase_change", "multiphase_two_atoms", "heatconduction_phase_change.lmp", cap=1, pre=[_atoms([[5, 5, 5], [5.6, 5, 5]], [1, 2])]),
    Shipped("two_atoms_phase_change", "multiphase_two_atoms", "phase_change.lmp", cap=1, pre=[_atoms([[5, 5, 5], [5.6, 5, 5]], [1, 2])]),
]


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_shipped_deck_through_the_shells(case, tmp_path):
    shim = shipped.build_shim()
    out = {}
    for who, exe, pre in (("ref", shipped.REF, None), ("b200", shipped.B200, shim)):
        wd = str(tmp_path / who)
        p = shipped.run_one(case, exe, wd, pre)
        assert p.returncode == 0 and "ERROR" not in p.stdout, who + ":\n" + p.stdout[-3000:] + p.stderr[-2000:]
        out[who] = (wd, p.stdout)
    assert "B200 engine: sph_oracle" in out["b200"][1]          # the shells ran, behind them the stand-in
    assert "B200 engine" not in out["ref"][1]
    shipped.compare_rows(shipped.thermo_block(out["ref"][1]), shipped.thermo_block(out["b200"][1]), case.tol, case.name + " thermo")
    files = list(case.files) + (["zz.dump"] if case.dump else [])
    for f in files:
        a = shipped.numeric_rows(os.path.join(out["ref"][0], f))
        b = shipped.numeric_rows(os.path.join(out["b200"][0], f))
        assert len(a) > 0, f
        shipped.compare_rows(a, b, case.tol, case.name + " " + f)
