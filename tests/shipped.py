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

The decks are read from /root/reference/examples/USER/sph at test time (this container only; the
tests skip where the reference is absent).  Two edits are applied to the text, both listed per deck
below: `run` lengths are capped so the CPU suite stays short, and a full-precision per-atom dump is
added in front of the first `run` so the final states can be compared digit by digit.
"""
import os
import re
import shutil
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLES = "/root/reference/examples/USER/sph"
REF = os.path.join(ROOT, "oracle", "_ref", "lmp_serial")
B200 = os.path.join(ROOT, "lammps-sph-multiphase_b200", "lammps", "_build", "lmp_b200")
BUILD = os.path.join(ROOT, "tests", "_build")
SHIM = os.path.join(BUILD, "liboracle_as_b200.so")

DUMP_COLS = "id type x y z vx vy vz fx fy fz c_zzrho c_zze"
DUMP_FMT = "%d %d " + " ".join(["%.17g"] * 11)


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
        with open(os.path.join(EXAMPLES, self.directory, self.deck)) as f:
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


def run_one(case, exe, workdir, preload=None):
    shutil.copytree(os.path.join(EXAMPLES, case.directory), workdir)
    for cmd in case.pre:
        subprocess.check_call(cmd, shell=True, cwd=workdir)
    with open(os.path.join(workdir, "zz_deck.lmp"), "w") as f:
        f.write(case.text())
    env = dict(os.environ)
    if preload:
        env["LD_PRELOAD"] = preload
    args = [exe] + (["-sf", "b200"] if preload else []) + ["-in", "zz_deck.lmp", "-log", "zz.log", "-echo", "none"] + case.var
    p = subprocess.run(args, cwd=workdir, capture_output=True, text=True, timeout=900, env=env)
    return p


NUM = re.compile(r"^[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)?$")


def numeric_rows(path):
    """every line of a text output as a list of tokens, numbers converted"""
    rows = []
    with open(path, errors="replace") as f:
        for l in f:
            rows.append([float(t) if NUM.match(t) else t for t in l.split()])
    return rows


def compare_rows(a, b, tol, what):
    """same line structure, same words, numbers to `tol` relative to the largest magnitude of their column block"""
    assert len(a) == len(b), "%s: %d lines against %d" % (what, len(a), len(b))
    worst = 0.0
    for i, (ra, rb) in enumerate(zip(a, b)):
        assert len(ra) == len(rb), "%s line %d: %r / %r" % (what, i + 1, ra, rb)
        na = np.array([v for v in ra if isinstance(v, float)])
        nb = np.array([v for v in rb if isinstance(v, float)])
        assert [v for v in ra if not isinstance(v, float)] == [v for v in rb if not isinstance(v, float)], "%s line %d: %r / %r" % (what, i + 1, ra, rb)
        if len(na):
            scale = max(np.abs(na).max(), 1e-300)
            worst = max(worst, float(np.abs(na - nb).max() / scale))
    assert worst <= tol, "%s: %.3g > %.3g" % (what, worst, tol)
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
