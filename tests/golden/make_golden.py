"""Generate tests/golden/<case>.npz from the REAL reference (oracle/_ref, built by
`make -C oracle ref` from /root/reference).  Runs only in the build container;
the fixtures it writes are committed and travel to the GPU box.

    python tests/golden/make_golden.py [case ...]

Per case the reference executes:  header + create + hot-path commands, then
    (state "init")  run 0  (state "s0" + full neighbor list)  run N  (state "sN")
All per-atom fields are stored for owned atoms in LAMMPS local order, full fp64.
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cases          # noqa: E402
import ref_lammps     # noqa: E402
from util import row_hashes  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
FULL_LIST_MAX = 20000   # store explicit (tag,image) lists only when this small


def generate(case):
    ref = ref_lammps.RefLammps()
    ref.command(case.header_text())
    ref.command(case.create)
    ref.command(case.lammps_text())
    mp = case.multiphase
    d = {}
    for k, v in ref.state(multiphase=mp).items():
        d["init_" + k] = v
    d["mass"] = ref.get("mass")
    ref.command("run 0")
    for k, v in ref.state(multiphase=mp).items():
        d["s0_" + k] = v
    num, jt, ji = ref.neighbor_list()
    d["nl_num"] = num; d["nl_hash"] = row_hashes(num, jt, ji)
    if len(jt) <= FULL_LIST_MAX:
        d["nl_jtag"] = jt; d["nl_jimage"] = ji
    d["s0_nghost"] = np.array(ref.nghost)
    shrink = any(c in case.boundary for c in "sm")
    if shrink:
        d["s0_box"] = np.array(ref.box()[:2])
    d["s0_virial"] = ref.virial()      # thermo prints at step 0 and at the last step: the pair virial is tallied there (integrate.cpp ev_set)
    cn = ref.cutneigh()
    d["cutneighsq"] = cn["cutneighsq"]; d["cutneighmax"] = np.array(cn["cutneighmax"]); d["cutghost"] = np.array(cn["cutghost"])
    d["neigh_params"] = np.array([cn["skin"], cn["every"], cn["delay"], cn["check"]])
    ref.command("run %d" % case.nsteps)
    for k, v in ref.state(multiphase=mp).items():
        d["sN_" + k] = v
    num, jt, ji = ref.neighbor_list()
    d["nlN_num"] = num; d["nlN_hash"] = row_hashes(num, jt, ji)
    d["sN_virial"] = ref.virial()
    if shrink:
        d["sN_box"] = np.array(ref.box()[:2])
    d["sN_nbuilds"] = np.array(ref.nbuilds); d["sN_ndanger"] = np.array(ref.ndanger)
    d["nsteps"] = np.array(case.nsteps)
    ref.close()
    path = os.path.join(OUT, case.name + ".npz")
    np.savez_compressed(path, **d)
    print("%-32s nlocal %6d -> %6d  nghost0 %6d  neigh/atom %.1f  builds %d  %.0f kB" % (
        case.name, len(d["init_type"]), len(d["sN_type"]), int(d["s0_nghost"]), d["nl_num"].mean(), int(d["sN_nbuilds"]),
        os.path.getsize(path) / 1024))


if __name__ == "__main__":
    names = sys.argv[1:] or list(cases.CASES)
    for n in names:
        generate(cases.CASES[n])
