"""Test harness: binds the CPU oracle (oracle/_build/libsph_oracle.so) through the same
ctypes ABI view the product uses, loads golden fixtures, and holds the comparison
recipe shared by the oracle-pinning tests (CPU) and the CUDA parity tests (GPU)."""
import ctypes
import importlib
import os
import subprocess
import numpy as np

import cases
from util import relerr, relerr_elem, row_hashes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = importlib.import_module("lammps-sph-multiphase_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libsph_oracle.so")

_oracle_api = None


def oracle_api():
    """the checker; built on demand with plain gcc (oracle/Makefile `port`)"""
    global _oracle_api
    if _oracle_api is None:
        src = os.path.join(ROOT, "oracle", "sph_oracle.c")
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"], stdout=subprocess.DEVNULL)
        _oracle_api = pkg._abi.bind(ctypes.CDLL(ORACLE_SO), "osph_")
    return _oracle_api


def oracle_sim(deck):
    return pkg.Sim(oracle_api(), deck)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def state_from(g, prefix, multiphase):
    keys = ["x", "v", "vest", "rho", "e", "cv", "type", "mask", "tag"]
    if multiphase:
        keys += ["rmass", "colorgradient"]
    return {k: g[prefix + k] for k in keys}


FIELDS_S0 = ("x", "v", "vest", "f", "rho", "drho", "e", "de")
FIELDS_MP = ("colorgradient", "rmass", "cv")


def compare_state(got, g, prefix, multiphase, tol, what):
    """every per-atom field within `tol` (max|a-b|/max|b|); ints exact"""
    errs = {}
    for k in ("type", "mask", "tag"):
        assert np.array_equal(got[k], g[prefix + k]), "%s: %s differs" % (what, k)
    for k in FIELDS_S0 + (FIELDS_MP if multiphase else ()):
        errs[k] = relerr(got[k], g[prefix + k])
    bad = {k: v for k, v in errs.items() if not (v <= tol)}
    assert not bad, "%s: fields beyond %g: %s" % (what, tol, bad)
    # the element-wise figure next to the norm-wise one (floor 1e-4 max|b|).  A component of size 1e-4 max still carries the
    # absolute rounding of the O(max) terms it is the sum of, so its budget is 1000x the norm-wise one (10x tighter than the
    # norm-wise bound alone implies for such a component)
    elem = {k: relerr_elem(got[k], g[prefix + k]) for k in errs}
    bad = {k: v for k, v in elem.items() if not (v <= 1e3 * tol)}
    assert not bad, "%s: element-wise errors beyond %g: %s" % (what, 1e3 * tol, bad)
    errs.update({k + "_elem": v for k, v in elem.items()})
    return errs


def compare_neighbors(sim, g, num_key="nl_num", hash_key="nl_hash"):
    """neighbor lists bit-exact: per-atom counts and order-independent row hashes of (tag,image)"""
    num, jt, ji = sim.neighbor_list()
    assert np.array_equal(num, g[num_key]), "numneigh differs for %d atoms" % int((num != g[num_key]).sum())
    assert np.array_equal(row_hashes(num, jt, ji), g[hash_key]), "neighbor rows differ"
    if num_key == "nl_num" and "nl_jtag" in g.files:
        assert np.array_equal(jt, g["nl_jtag"]) and np.array_equal(ji, g["nl_jimage"])
    return int(num.sum())


def run_case(make_sim, name, tol_step=1e-10, tol_traj=None):
    """the reference's own sequence: (init) run 0 (s0) run N (sN)"""
    case = cases.CASES[name]
    g = load_golden(name)
    deck = case.deck()
    # host mirror tables == the reference's Pair/Neighbor::init output, bit for bit
    assert np.array_equal(deck.cutneighsq, g["cutneighsq"])
    assert deck.cutneighmax == float(g["cutneighmax"]) and deck.cutghost == float(g["cutghost"])
    assert [deck.skin, deck.every, deck.delay, deck.check] == list(g["neigh_params"])
    if not case.multiphase:
        assert np.array_equal(deck.mass_[1:], g["mass"][1:])
    sim = make_sim(deck)
    sim.set_atoms(**state_from(g, "init_", case.multiphase))
    sim.request_virial()          # thermo output at step 0 and at the last step: the reference tallies the pair virial there
    sim.setup()
    e0 = compare_state(sim.get_atoms(), g, "s0_", case.multiphase, tol_step, name + " run 0")
    e0["virial"] = relerr(sim.virial(), g["s0_virial"])
    assert e0["virial"] <= 10 * tol_step, "%s run 0: pair virial off by %g" % (name, e0["virial"])
    compare_neighbors(sim, g)
    assert sim.natoms()[1] == int(g["s0_nghost"]), "ghost count"
    if "s0_box" in g:             # boundary s / m: the box Domain::reset_box fitted, bit for bit
        assert np.array_equal(np.array(sim.box()), g["s0_box"]), "%s run 0: shrink-wrapped box %s vs %s" % (name, sim.box(), g["s0_box"])
    sim.setup()
    sim.request_virial()
    sim.run(case.nsteps)
    got = sim.get_atoms()
    assert len(got["type"]) == len(g["sN_type"]), "particle count after run"
    eN = compare_state(got, g, "sN_", case.multiphase, tol_traj or case.tol_traj, name + " run N")
    eN["virial"] = relerr(sim.virial(), g["sN_virial"])
    assert eN["virial"] <= 10 * (tol_traj or case.tol_traj), "%s run N: pair virial off by %g" % (name, eN["virial"])
    compare_neighbors(sim, g, "nlN_num", "nlN_hash")
    if "sN_box" in g:
        assert relerr(np.array(sim.box()), g["sN_box"]) <= (tol_traj or case.tol_traj), "%s run N: shrink-wrapped box" % name
    c = sim.counters()
    assert c["builds"] == int(g["sN_nbuilds"]), "neighbor builds %d vs reference %d" % (c["builds"], int(g["sN_nbuilds"]))
    sim.close()
    return e0, eN
