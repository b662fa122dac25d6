"""CPU: the host mirror of the reference's command parsing (lammps-sph-multiphase_b200/deck.py) --
argument counts, keywords and error messages of the styles added for SURVEY 8(f) follow the reference
(pair_sph_idealgas.cpp:210-238, fix_setforce.cpp:40-110, fix_setmesode.cpp:38-78, fix_dt_reset.cpp:40-98)."""
import importlib

import numpy as np
import pytest

pkg = importlib.import_module("lammps-sph-multiphase_b200")
Deck, DeckError = pkg.deck.Deck, pkg.deck.DeckError


def _deck(ntypes=2):
    d = Deck(dimension=3, boundary="p p p", box=((0, 0, 0), (1, 1, 1)), atom_style="meso", ntypes=ntypes, units="lj")
    d.mass("*", 1.0)
    return d


def test_idealgas_coeff_and_unmirrored_viscosity():
    d = _deck()
    d.pair_style("hybrid/overlay", "sph/rhosum 1", "sph/idealgas")
    d.pair_coeff("* *", "sph/rhosum", 0.1)
    with pytest.raises(DeckError, match="Incorrect args"):
        d.pair_coeff("* *", "sph/idealgas", 0.75)                 # needs viscosity and h
    d.pair_coeff("* *", "sph/idealgas", 0.75, 0.1)
    d.neighbor(0.01); d.timestep(0.01); d.fix("i", "all", "meso")
    d.init()
    gas = [s for s in d.styles if s.name == "sph/idealgas"][0]
    # PairSPHIdealGas::init_one mirrors only cut (pair_sph_idealgas.cpp:244-253)
    assert gas.cut[2, 1] == gas.cut[1, 2] == 0.1
    assert gas.viscosity[1, 2] == 0.75 and gas.viscosity[2, 1] == 0.0


def test_fix_setforce_setmesode_dt_reset_arguments():
    d = _deck(1)
    d.region("r", "block", 0.2, 0.4, "EDGE", "EDGE", "EDGE", "EDGE")
    d.fix("a", "all", "setforce", "NULL", 0.0, 0.0)
    with pytest.raises(DeckError, match="Illegal fix setforce"):
        d.fix("b", "all", "setforce", 0.0, 0.0)
    with pytest.raises(DeckError, match="constant values"):
        d.fix("b", "all", "setforce", "v_fx", 0.0, 0.0)
    d.fix("c", "all", "setmesode", 0.5, "region", "r")
    with pytest.raises(DeckError, match="does not exist"):
        d.fix("c2", "all", "setmesode", 0.5, "region", "nope")
    d.fix("e", "all", "dt/reset", 1, "NULL", 1e-4, 5e-4, "units", "box")
    with pytest.raises(DeckError, match="Illegal fix dt/reset"):
        d.fix("e2", "all", "dt/reset", 0, "NULL", 1e-4, 5e-4, "units", "box")
    with pytest.raises(DeckError, match="Illegal fix dt/reset"):
        d.fix("e3", "all", "dt/reset", 1, 2e-4, 1e-4, 5e-4, "units", "box")     # tmin >= tmax
    with pytest.raises(DeckError, match="units box"):
        d.fix("e4", "all", "dt/reset", 1, "NULL", 1e-4, 5e-4)
    kinds = [f[0] for f in d.fixes]
    assert kinds == ["setforce", "setmesode", "dt/reset"]
    assert d.fixes[0][2] == ([0, 1, 1], [0.0, 0.0, 0.0])


def test_variable_formulas_engine_compiler_vs_oracle():
    """the engine's formula compiler + postfix evaluator (csrc/b200_expr.cuh, host side) against the oracle's direct evaluator
    (Variable::evaluate restated) and against values worked out by hand; precedence and left association as variable.cpp:99-107,1641"""
    import ctypes as C
    import harness
    eng, ora = pkg.load(), harness.oracle_api()
    rng = np.random.default_rng(5)
    formulas = ["mass*-9.81", "mass*0.5*((y<0.2)-(y>0.2))", "2^3^2", "-2^2", "1-2-3", "2*3%4", "1.5e-3*x+y/z", "!(x>y)||(vx<=vy&&fz!=0)",
                "sqrt(abs(fx))+exp(-x)*ln(mass)+log(100)", "atan2(y,x)+sin(PI/2)+round(-2.5)+ceil(x)+floor(y)", "(type==2)*id+step*dt",
                "((3.0-1)*y/1.0-((3.0-1)*0.1-1.0)/1.0)*(y-0.1<1.0)+3.0*(y-0.1>1.0)", "1.0+sqrt(x)*(x<=0.3)+abs(-0.5)*(x>0.3)", "2--3" if False else "2-(-3)",
                # what FixGravityB200 composes for `fix gravity v_gmag vector v_gx 1 0` (fix_b200.cpp): massone * (magnitude * xdir / length)
                "mass*((-9.81*(step>3)*(1.0+0.01*step))*((0.02*step*dt/1.0e-4)/sqrt((0.02*step*dt/1.0e-4)*(0.02*step*dt/1.0e-4)+(1)*(1))))",
                "mass*((-9.81*(step>3)*(1.0+0.01*step))*((1)/sqrt((0.02*step*dt/1.0e-4)*(0.02*step*dt/1.0e-4)+(1)*(1))))", "mass*((-9.81)*(0.0))",
                # what FixAddForceB200 composes for `every N` / `region ID` (block with EDGE bounds, sphere)
                "(2.5)*((step%2)==0)*((x>=(-1e+20))&&(x<=(1.5))&&(y>=(0.10000000000000001))&&(y<=(1e+20))&&(z>=(-1e+20))&&(z<=(1e+20)))",
                "(mass*-9.81)*((step%17)==0)*(sqrt((x-(0.5))*(x-(0.5))+(y-(0.5))*(y-(0.5))+(z-(0))*(z-(0)))<=(0.75))"]
    known = {"2^3^2": 64.0, "-2^2": 4.0, "1-2-3": -4.0, "2*3%4": 2.0, "2-(-3)": 5.0}
    for f in formulas:
        for _ in range(5):
            atom = np.ascontiguousarray(rng.uniform(0.05, 2.0, 12)); ty, tag, step, dt = int(rng.integers(1, 4)), int(rng.integers(1, 1000)), 17.0, 2.5e-4
            a, b = C.c_double(), C.c_double()
            assert eng.formula_check(f.encode(), atom.ctypes.data_as(C.POINTER(C.c_double)), ty, tag, step, dt, C.byref(a)) == 0, (f, eng.last_error())
            assert ora.formula_check(f.encode(), atom.ctypes.data_as(C.POINTER(C.c_double)), ty, tag, step, dt, C.byref(b)) == 0, f
            assert a.value == b.value or abs(a.value - b.value) <= 1e-15 * abs(b.value), (f, a.value, b.value)
            if f in known:
                assert a.value == known[f], (f, a.value)
    for bad in ("mass*", "foo(x)", "c_pe*2", "x+", "random(0,1,5)", "(x"):
        assert eng.formula_check(bad.encode(), None, 1, 1, 0.0, 0.0, None) < 0, bad


def test_addforce_and_variable_parsing():
    """fix addforce fx fy fz with v_name arguments (fix_addforce.cpp:40-110) and variable references (variable.cpp): what the mirror accepts and refuses"""
    d = _deck()
    d.variable("gx", "equal", "0.5*2.0")
    d.variable("bodyfx", "atom", "mass*v_gx*((y<0.2)-(y>0.2))")
    d.fix("f1", "all", "addforce", "v_bodyfx", 0.0, "v_gx")
    style, bit, (vals, forms) = d.fixes[-1]
    assert style == "addforce" and vals == [0.0, 0.0, 0.0]
    assert forms == ["mass*(0.5*2.0)*((y<0.2)-(y>0.2))", None, "0.5*2.0"]
    with pytest.raises(DeckError):
        d.fix("f2", "all", "addforce", "v_nope", 0.0, 0.0)                     # unknown variable
    with pytest.raises(DeckError):
        d.fix("f3", "all", "addforce", 1.0, 0.0, 0.0, "every", 2)             # keywords are not taken
    d.variable("a", "equal", "v_b+1"); d.variable("b", "equal", "v_a+1")
    with pytest.raises(DeckError):
        d.fix("f4", "all", "addforce", "v_a", 0.0, 0.0)                        # circular
    with pytest.raises(DeckError):
        d.variable("s", "string", "abc")                                       # only equal / atom styles
    d.fix("f5", "all", "setmeso", "meso_e", "v_gx")
    assert d.fixes[-1][0] == "setmeso/var" and d.fixes[-1][2][1] == "0.5*2.0"
