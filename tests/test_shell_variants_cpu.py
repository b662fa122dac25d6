"""CPU: what the USER-B200 shells refuse, and deck idioms beyond the shipped examples, through `lmp_b200 -sf b200` with the oracle
behind the C-ABI (tests/shipped.py) against `lmp_serial`.  The base text is the shipped water_collapse deck (read from the reference at
test time) with one edit per case."""
import os
import re

import pytest

import shipped
from shipped import Shipped

pytestmark = pytest.mark.skipif(not shipped.available(), reason="needs /root/reference, oracle/_ref/lmp_serial and lmp_b200")

REFUSED = [
    # (edit: regex -> replacement on water_collapse.lmp, message the shell must print)
    ("per_atom_virial", (r"^thermo_style.*$", "compute st all stress/atom NULL\ncompute sts all reduce sum c_st[1]\nthermo_style custom step c_sts"),
     "per-atom energy / virial tallies"),
    ("host_stepping_fix", (r"^fix\s+2d_fix.*$", "fix 2d_fix all enforce2d\nfix wl water wall/reflect ylo EDGE"), "has no /b200 variant"),
    ("end_of_step_fix_that_writes", (r"^fix\s+2d_fix.*$", "fix 2d_fix all enforce2d\nfix mom water momentum 1 linear 1 1 0"), "has no /b200 variant"),
    ("neigh_exclude", (r"^neigh_modify.*$", "neigh_modify every 5 delay 0 check no exclude type 1 2"), "neigh_modify exclude is not supported"),
    ("neighbor_nsq", (r"^neighbor\s.*$", "neighbor ${skin} nsq"), "supports neighbor style bin"),
    ("compute_with_neighbor_list", (r"^thermo_style.*$", "compute rd all rdf 20\nfix rdav all ave/time 5 1 5 c_rd file zz.rdf mode vector\nthermo_style custom step ke"),
     "needs a host neighbor list"),
    ("comm_mode_multi", (r"^neigh_modify.*$", "neigh_modify every 5 delay 0 check no\ncomm_modify mode multi"), "comm_modify mode single"),
    ("atom_leaves_a_fixed_face", (r"^run\s+\S+.*$", "group one id 9000\nset group one y 7.9995\nvelocity one set 0.0 100.0 0.0 units box\nthermo 1\nrun 8"), "beyond a fixed box face"),
    ("region_style", (r"^fix\s+2d_fix.*$", "fix 2d_fix all enforce2d\nregion cyl cylinder z 1.0 1.0 0.5 EDGE EDGE units box\nfix sm water setmeso meso_e 0.1 region cyl"),
     "supports block and sphere regions"),
    ("addforce_energy_keyword", (r"^fix\s+2d_fix.*$", "fix 2d_fix all enforce2d\nvariable en atom 0.0\nfix af water addforce 0.0 1.0 0.0 energy v_en"), "does not take the energy keyword"),
    ("setforce_variable", (r"^fix\s+2d_fix.*$", "fix 2d_fix all enforce2d\nvariable fz equal 0.0\nfix sf bc setforce 0.0 0.0 v_fz"), "constant values"),
    ("newton_off", (r"^newton\s+on", "newton off"), "requires newton on"),
    ("variable_gravity_angle", (r"^fix\s+gfix.*$", "variable ang equal 10.0+0.01*step\nfix gfix water gravity 9.81 chute v_ang"), "supports variables for the magnitude"),
]


@pytest.mark.parametrize("name,edit,message", REFUSED, ids=[r[0] for r in REFUSED])
def test_refused_with_a_message(name, edit, message, tmp_path):
    """no silent fallback and no crash: the run stops with the shell's own message before any step is taken"""
    case = Shipped(name, "water_collapse", "water_collapse.lmp", cap=10, subs=[edit], dump=False)
    p = shipped.run_one(case, shipped.B200, str(tmp_path / "b200"), shipped.build_shim())
    assert p.returncode == 1, (p.returncode, p.stdout[-1500:], p.stderr[-500:])
    assert "ERROR" in p.stdout and message in p.stdout, p.stdout[-1500:]
    assert "Loop time" not in p.stdout or name == "atom_leaves_a_fixed_face"      # (that one stops inside the run)


VARIANTS = [
    # run N upto / start-stop keywords, runs that continue an earlier one (their setups see vest != v)
    ("run_keywords", [(r"^run\s+\S+.*$", "run 12\nrun 30 upto\nrun 10 start 0 stop 200")], 1e-9),
    # restart files written during a run and read back into a fresh deck would need a second deck: here the periodic write itself (an output step)
    ("restart_every", [(r"^run\s+\S+.*$", "restart 8 zz.restart\nrun 24")], 1e-9),
    # thermo every step (every step is an output step: one-step segments), dump every 3
    ("every_step_output", [(r"^thermo\s+10", "thermo 1"), (r"^run\s+\S+.*$", "run 9")], 1e-9),
    # the host changes the system between runs: atoms deleted and displaced, fields set, a fix removed and another added, timestep and
    # neighbor settings changed -- every setup uploads the host arrays and re-registers the deck
    ("edits_between_runs", [(r"^run\s+\S+.*$", "run 8\nregion cut block 0.5 0.9 0.3 0.6 EDGE EDGE units box\ndelete_atoms region cut\n"
                                                "displace_atoms water move 0.0 0.002 0.0 units box\nset group water meso_e 0.5\nunfix dtfix\ntimestep 2.0e-5\n"
                                                "neigh_modify every 3 delay 0 check yes\nfix frz bc setforce 0.0 NULL 0.0\nthermo_style custom step ke c_esph press\nrun 10")], 1e-9),
    # fix gravity with equal-style variables (magnitude switched on over the first steps, a direction that tilts with time): the formula
    # FixGravity::post_force evaluates goes to the engine's per-atom evaluator; f_gfix on the thermo line follows the same variables
    ("variable_gravity", [(r"^fix\s+gfix.*$", "variable gmag equal -9.81*(step>3)*(1.0+0.01*step)\nvariable gx equal 0.02*step*dt/1.0e-4\n"
                                              "fix gfix water gravity v_gmag vector v_gx 1 0"), (r"^run\s+\S+.*$", "run 20")], 1e-9),
    # run N every M "command" with the default pre yes: the command writes per-atom data on the host, every chunk sets up again (uploads)
    ("run_every_writes", [(r"^run\s+\S+.*$", 'run 12 every 4 "velocity water scale 0.5" "set group water meso_e 0.1"')], 1e-9),
    # fix addforce with `every N` and `region ID` (block, sphere), constant and atom-style variable components
    ("addforce_every_region", [(r"^fix\s+2d_fix.*$", "fix 2d_fix all enforce2d\nregion rb block 0.5 1.5 0.2 1.0 EDGE EDGE units box\nregion rs sphere 1.0 0.5 0.0 0.4 units box\n"
                                                      "variable push atom mass*2.0*(y>0.3)\nfix a1 water addforce v_push 0.5e-3 0.0 every 3 region rb\n"
                                                      "fix a2 water addforce 0.0 -1.0e-3 0.0 region rs\nfix a3 water addforce 1.0e-4 0.0 0.0 every 2"),
                               (r"^run\s+\S+.*$", "run 14")], 1e-9),
    # fix ave/time over a compute reduce, fix ave/atom of a per-atom compute (END_OF_STEP, read-only, evaluated on their own steps)
    ("fix_ave", [(r"^run\s+\S+.*$", "fix avt all ave/time 2 3 6 c_esph file zz.avt\nfix ava all ave/atom 1 4 4 c_rho_peratom\nrun 12")], 1e-9),
]


@pytest.mark.parametrize("name,edits,tol", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_deck_idioms_match_the_reference(name, edits, tol, tmp_path):
    case = Shipped(name, "water_collapse", "water_collapse.lmp", cap=10 ** 9, subs=edits, files=["zz.avt"] if name == "fix_ave" else [])
    case.cap = 4           # only the interval of the appended full-precision dump (text() caps literal run lengths at 10**9 = not at all)
    text = case.text
    case.text = lambda: re.sub(r"^(\s*)run (\d+)", lambda m: "%srun %s" % (m.group(1), m.group(2)), text(), flags=re.M)
    out = {}
    for who, exe, pre in (("ref", shipped.REF, None), ("b200", shipped.B200, shipped.build_shim())):
        wd = str(tmp_path / who)
        p = shipped.run_one(case, exe, wd, pre)
        assert p.returncode == 0 and "ERROR" not in p.stdout, who + ":\n" + p.stdout[-3000:] + p.stderr[-2000:]
        out[who] = (wd, p.stdout)
    assert "B200 engine: sph_oracle" in out["b200"][1]
    ta, tb = shipped.thermo_block(out["ref"][1]), shipped.thermo_block(out["b200"][1])
    assert len(ta) > 4
    shipped.compare_rows(ta, tb, tol, name + " thermo")
    for f in list(case.files) + ["zz.dump"]:
        a = shipped.numeric_rows(os.path.join(out["ref"][0], f))
        b = shipped.numeric_rows(os.path.join(out["b200"][0], f))
        assert len(a) > 0, f
        shipped.compare_rows(a, b, tol, name + " " + f, shipped.DUMP_VECTORS if f == "zz.dump" else ())


def test_phase_change_keeps_its_state_across_runs(tmp_path):
    """bubble_on_wall.lmp with its long `run ... pre no every` replaced by three plain runs: every `run` sets up again and re-registers
    the deck, fix phase_change must go on with its next step and its RanPark stream (fix_phase_change.cpp:116,345) -- atom counts,
    the deck's own rg.dat (fix print 1) and the final dump against lmp_serial"""
    case = Shipped("bubble_on_wall_three_runs", "bubble_on_wall", "bubble.lmp", var=["-var", "dname", "data"], cap=10 ** 9, files=["data/rg.dat"],
                   subs=[(r"^run\s+10000000 pre no\s+post no every 1000 &\n.*&\n.*$", "run 15\nrun 15\nrun 15")])
    case.cap = 15
    out = {}
    for who, exe, pre in (("ref", shipped.REF, None), ("b200", shipped.B200, shipped.build_shim())):
        wd = str(tmp_path / who)
        p = shipped.run_one(case, exe, wd, pre)
        assert p.returncode == 0 and "ERROR" not in p.stdout, who + ":\n" + p.stdout[-3000:] + p.stderr[-2000:]
        out[who] = (wd, p.stdout)
    ta, tb = shipped.thermo_block(out["ref"][1]), shipped.thermo_block(out["b200"][1])
    natoms = [r[1] for r in ta if r[0] == "atoms"]
    assert len(natoms) == 4 and natoms[-1] > natoms[0], natoms          # run 0 + three runs; atoms were inserted
    shipped.compare_rows(ta, tb, 1e-9, "thermo")
    for f in ("data/rg.dat", "zz.dump"):
        shipped.compare_rows(shipped.numeric_rows(os.path.join(out["ref"][0], f)), shipped.numeric_rows(os.path.join(out["b200"][0], f)), 1e-9, f, shipped.DUMP_VECTORS if f == "zz.dump" else ())


@pytest.mark.parametrize("form", ["ENERGY 0.02", "${prob} attempt 3"], ids=["energy_rate", "attempt"])
def test_phase_change_argument_forms(form, tmp_path):
    """fix phase_change's other argument forms (fix_phase_change.cpp:57-79,358-390: `ENERGY rate` instead of a probability, `attempt N`)
    through the shell's own parser, on the shipped bubble_on_wall deck"""
    case = Shipped("bubble_on_wall_" + form.split()[0], "bubble_on_wall", "bubble.lmp", var=["-var", "dname", "data"], cap=10 ** 9, files=["data/rg.dat"],
                   subs=[(r"12345 \$\{prob\} region rflow", "12345 " + form + " region rflow"),
                         (r"^run\s+10000000 pre no\s+post no every 1000 &\n.*&\n.*$", "run 40")])
    case.cap = 40
    out = {}
    for who, exe, pre in (("ref", shipped.REF, None), ("b200", shipped.B200, shipped.build_shim())):
        wd = str(tmp_path / who)
        p = shipped.run_one(case, exe, wd, pre)
        assert p.returncode == 0 and "ERROR" not in p.stdout, who + ":\n" + p.stdout[-3000:] + p.stderr[-2000:]
        out[who] = (wd, p.stdout)
    ta, tb = shipped.thermo_block(out["ref"][1]), shipped.thermo_block(out["b200"][1])
    shipped.compare_rows(ta, tb, 1e-9, "thermo")
    for f in ("data/rg.dat", "zz.dump"):
        shipped.compare_rows(shipped.numeric_rows(os.path.join(out["ref"][0], f)), shipped.numeric_rows(os.path.join(out["b200"][0], f)), 1e-9, f,
                             shipped.DUMP_VECTORS if f == "zz.dump" else ())


HEAT_VARIANTS = [
    # regions are in lattice units in this deck (no `units box`): block / sphere, region / noregion, constant and variable values
    ("setmesode_region", "region hot block 20 30 EDGE EDGE EDGE EDGE\nfix sde all setmesode 0.002 region hot"),
    ("setmeso_sphere_noregion", "region ball sphere 70 5 0 8\nfix sm all setmeso meso_e 1.5 noregion ball"),
    ("setmeso_equal_variable", "variable ramp equal 1.0+0.001*step\nfix sm all setmeso meso_e v_ramp region left"),
    ("setmeso_atom_variable_temperature", "variable prof atom 1.0+0.01*x\nfix sm all setmeso meso_t v_prof region right"),
]


@pytest.mark.parametrize("name,lines", HEAT_VARIANTS, ids=[v[0] for v in HEAT_VARIANTS])
def test_setmeso_forms_on_the_heat_deck(name, lines, tmp_path):
    """fix setmeso / setmesode argument forms (fix_setmeso.cpp:38-89,180-271, fix_setmesode.cpp:38-78) added to the shipped 2-D heat deck"""
    case = Shipped(name, "heatconduction", "sph_heat_conduction_2d.lmp", cap=40, subs=[(r"^fix\s+integrate_fix.*$", "fix integrate_fix all meso/stationary\n" + lines)])
    out = {}
    for who, exe, pre in (("ref", shipped.REF, None), ("b200", shipped.B200, shipped.build_shim())):
        wd = str(tmp_path / who)
        p = shipped.run_one(case, exe, wd, pre)
        assert p.returncode == 0 and "ERROR" not in p.stdout, who + ":\n" + p.stdout[-3000:] + p.stderr[-2000:]
        out[who] = (wd, p.stdout)
    shipped.compare_rows(shipped.thermo_block(out["ref"][1]), shipped.thermo_block(out["b200"][1]), 1e-10, "thermo")
    a, b = (shipped.numeric_rows(os.path.join(out[w][0], "zz.dump")) for w in ("ref", "b200"))
    shipped.compare_rows(a, b, 1e-10, "zz.dump", shipped.DUMP_VECTORS)
    assert len({r[12] for r in a if len(r) == 13}) > 5          # the energies did change
