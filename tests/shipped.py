"""CPU check of the USER-B200 shells on the reference's OWN example decks (test infrastructure).

`lmp_b200` is the reference LAMMPS + the C++ shells of lammps/USER-B200 + `libb200sph.so`.  Here the
library behind the C-ABI is swapped -- by LD_PRELOAD, for this test only -- for the CPU oracle with
its `osph_*` exports renamed to `b200_*` (objcopy on the oracle's object file, tests/_build/).  What
runs is therefore: shipped deck text -> reference parser -> /b200 shells (VerletB200 segmenting,
fix / pair shells, download on output steps) -> C-ABI -> oracle, and it is compared with the
unmodified `lmp_serial` on the same text.  It pins the HOST side of the drop-in (the shells accept
every command of the shipped decks and hand the library what the reference classes would have
computed with); the CUDA engine behind the same ABI is pinned against the oracle by the -m gpu tests.
The product never loads this shim: `libb200sph.so` is resolved through lmp_b200's rpath unless this
test preloads the stand-in.

The decks are read from /root/reference/examples/USER/sph at test time (the CPU tests skip where the
reference is absent; tests/test_gpu_zz_shipped.py, which runs the same table on the CUDA engine, takes
them from the archive `make -C oracle ref` packs beside the reference binaries).  Two edits are applied
to the text, both listed per deck below: `run` lengths are capped so the suites stay short, and a
full-precision per-atom dump is added in front of the first `run` so the final states can be compared
digit by digit.  Outputs are compared column by column (compare_rows): words, ids, types, steps and counts
exactly, every numeric column against its own largest magnitude, vector components against the vector's.
"""
import os
import re
import shutil
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLES = "/root/reference/examples/USER/sph"
STAGED = os.path.join(ROOT, "oracle", "_ref", "examples_sph.tar.gz")     # `make -C oracle ref` packs the example inputs next to the reference binaries
REF = os.path.join(ROOT, "oracle", "_ref", "lmp_serial")
B200 = os.path.join(ROOT, "lammps-sph-multiphase_b200", "lammps", "_build", "lmp_b200")
BUILD = os.path.join(ROOT, "tests", "_build")
SHIM = os.path.join(BUILD, "liboracle_as_b200.so")

DUMP_COLS = "id type x y z vx vy vz fx fy fz c_zzrho c_zze"
DUMP_FMT = "%d %d " + " ".join(["%.17g"] * 11)


_examples = None


def examples_dir():
    """the reference's examples/USER/sph: in place where /root/reference exists, else unpacked (once per process) from the archive that
    oracle/Makefile staged beside the reference binaries -- the GPU box has no /root/reference.  None if neither is there."""
    global _examples
    if _examples is None:
        if os.path.isdir(EXAMPLES):
            _examples = EXAMPLES
        elif os.path.exists(STAGED):
            import tarfile
            import tempfile
            d = tempfile.mkdtemp(prefix="sph_examples_")
            with tarfile.open(STAGED) as tf:
                tf.extractall(d, filter="tar")
            _examples = os.path.join(d, "sph")
        else:
            _examples = ""
    return _examples or None


def available():
    return os.path.isdir(EXAMPLES) and os.path.exists(REF) and os.path.exists(B200) and shutil.which("objcopy") is not None


def build_shim():
    """oracle/sph_oracle.c -> object -> every osph_X that has a `#define b200_X osph_X` line in
    oracle/sph_oracle.h renamed back to b200_X -> tests/_build/liboracle_as_b200.so"""
    src = os.path.join(ROOT, "oracle", "sph_oracle.c")
    hdr = os.path.join(ROOT, "oracle", "sph_oracle.h")
    if os.path.exists(SHIM) and os.path.getmtime(SHIM) > max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(os.path.join(ROOT, "include", "b200_sph.h"))):
        return SHIM
    os.makedirs(BUILD, exist_ok=True)
    obj = os.path.join(BUILD, "sph_oracle.o")
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c99", "-ffp-contract=off", "-c", src, "-o", obj])
    syms = os.path.join(BUILD, "rename.txt")
    with open(hdr) as f, open(syms, "w") as g:
        for m in re.finditer(r"^#define\s+(b200_\w+)\s+(osph_\w+)", f.read(), re.M):
            g.write("%s %s\n" % (m.group(2), m.group(1)))
    subprocess.check_call(["objcopy", "--redefine-syms=" + syms, obj, obj + "2"])
    subprocess.check_call(["gcc", "-shared", "-o", SHIM, obj + "2", "-lm"])
    return SHIM


class Shipped:
    """one shipped example: directory, main deck, -var arguments of its run.sh (sizes reduced where the deck takes them
    as variables), the cap on every `run`, extra text substitutions (regex -> replacement), output files to compare"""

    def __init__(self, name, directory, deck, var=(), cap=40, subs=(), files=(), tol=1e-9, dump=True, pre=()):
        self.name, self.directory, self.deck, self.var, self.cap = name, directory, deck, list(var), cap
        self.subs, self.files, self.tol, self.dump, self.pre = list(subs), list(files), tol, dump, list(pre)

    def text(self):
        with open(os.path.join(examples_dir(), self.directory, self.deck)) as f:
            t = f.read()
        for pat, rep in self.subs:
            t, n = re.subn(pat, rep, t, flags=re.M)
            assert n, (self.name, pat)

        def cap(m):
            arg = m.group(2)
            try:
                n = min(int(arg), self.cap)
            except ValueError:
                n = self.cap
            return "%srun %d" % (m.group(1), n)
        t = re.sub(r"^(\s*)run\s+(\S+)", cap, t, flags=re.M)
        if self.dump:
            extra = "\n".join(["compute zzrho all meso_rho/atom", "compute zze all meso_e/atom",
                               "dump zzfin all custom %d zz.dump %s" % (self.cap, DUMP_COLS),
                               'dump_modify zzfin sort id format "%s"' % DUMP_FMT,
                               "thermo_modify format float %.15g", ""])
            m = re.search(r"^\s*run\s", t, flags=re.M)
            t = t[:m.start()] + extra + t[m.start():]
        return t


def run_one(case, exe, workdir, preload=None, suffix=False, timeout=900):
    shutil.copytree(os.path.join(examples_dir(), case.directory), workdir)
    for cmd in case.pre:
        subprocess.check_call(cmd, shell=True, cwd=workdir)
    with open(os.path.join(workdir, "zz_deck.lmp"), "w") as f:
        f.write(case.text())
    env = dict(os.environ)
    if preload:
        env["LD_PRELOAD"] = preload
    args = [exe] + (["-sf", "b200"] if (preload or suffix) else []) + ["-in", "zz_deck.lmp", "-log", "zz.log", "-echo", "none"] + case.var
    p = subprocess.run(args, cwd=workdir, capture_output=True, text=True, timeout=timeout, env=env)
    return p


NUM = re.compile(r"^[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)?$")


def numeric_rows(path):
    """every line of a text output as a list of tokens, numbers converted"""
    rows = []
    with open(path, errors="replace") as f:
        for l in f:
            rows.append([float(t) if NUM.match(t) else t for t in l.split()])
    return rows


DUMP_VECTORS = [(2, 3, 4), (5, 6, 7), (8, 9, 10)]       # x, v, f of the appended dump: the components of a vector share one scale


def compare_rows(a, b, tol, what, vectors=()):
    """same line structure and the same words; numbers column by column: |a - b| <= tol * (largest |a| of that column over all lines of
    the same shape), so ids, types, steps and counts must agree exactly and every field is measured against its own scale.
    `vectors`: tuples of token positions whose columns are the components of one vector field (scaled by its largest component, as
    tests/harness.py measures per-atom vectors -- a component that is zero by symmetry holds only rounding noise)"""
    assert len(a) == len(b), "%s: %d lines against %d" % (what, len(a), len(b))
    groups = {}
    for i, (ra, rb) in enumerate(zip(a, b)):
        assert len(ra) == len(rb), "%s line %d: %r / %r" % (what, i + 1, ra, rb)
        wa = tuple(v if not isinstance(v, float) else None for v in ra)
        wb = tuple(v if not isinstance(v, float) else None for v in rb)
        assert wa == wb, "%s line %d: %r / %r" % (what, i + 1, ra, rb)
        groups.setdefault(wa, []).append(i)
    worst = 0.0
    for shape, rows in groups.items():
        cols = [c for c, w in enumerate(shape) if w is None]
        if not cols:
            continue
        A = np.array([[a[i][c] for c in cols] for i in rows]); B = np.array([[b[i][c] for c in cols] for i in rows])
        scale = np.maximum(np.abs(A).max(axis=0), 1e-300)
        for vec in vectors:
            idx = [cols.index(c) for c in vec if c in cols]
            if len(idx) == len(vec):
                scale[idx] = scale[idx].max()
        err = np.abs(A - B) / scale
        k = np.unravel_index(err.argmax(), err.shape)
        assert err[k] <= tol, "%s line %d column %d: %r against %r (%.3g > %.3g of the column's scale %.3g)" % (
            what, rows[k[0]] + 1, cols[k[1]] + 1, A[k], B[k], err[k], tol, scale[k[1]])
        worst = max(worst, float(err.max()))
    return worst


def thermo_block(log):
    """lines of the log between `Step ...` headers and `Loop time` (the thermo output of every run), plus the run statistics that
    depend on the engine (atom count, neighbor list builds)"""
    rows, on = [], False
    for l in log.splitlines():
        if l.startswith("Step "):
            on = True
        elif l.startswith("Loop time"):
            on = False
            rows.append(["atoms"] + l.split()[-2:-1])
        elif l.startswith("Neighbor list builds"):
            rows.append(l.split())
        if on:
            rows.append(l.split())
    return [[float(t) if NUM.match(t) else t for t in r] for r in rows]


# ---- the decks -------------------------------------------------------------------------------------------------------------------
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
