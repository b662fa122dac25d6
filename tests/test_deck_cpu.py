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
