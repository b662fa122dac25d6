"""ref_lammps.py -- TEST INFRASTRUCTURE ONLY.

ctypes driver of the unmodified reference build (oracle/_ref/liblammps_ref.so via
oracle/_ref/librefshim.so; recipe: oracle/Makefile).  Used in this container to
pin the oracle and to generate tests/golden/*.npz (tests/golden/make_golden.py),
and by bench.py --impl reference.  Never imported by the product package.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, "_ref", "librefshim.so")


def available():
    return os.path.exists(SHIM)


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(SHIM)
        _lib.refshim_open.restype = C.c_void_p
        _lib.refshim_open.argtypes = [C.c_int]
        for n in ("close", "command", "file"):
            getattr(_lib, "refshim_" + n).restype = None
        _lib.refshim_close.argtypes = [C.c_void_p]
        _lib.refshim_command.argtypes = [C.c_void_p, C.c_char_p]
        _lib.refshim_file.argtypes = [C.c_void_p, C.c_char_p]
        for n in ("nlocal", "nghost", "ntypes"):
            f = getattr(_lib, "refshim_" + n); f.restype = C.c_int; f.argtypes = [C.c_void_p]
        for n in ("ntimestep", "nbuilds", "ndanger"):
            f = getattr(_lib, "refshim_" + n); f.restype = C.c_longlong; f.argtypes = [C.c_void_p]
        _lib.refshim_dt.restype = C.c_double; _lib.refshim_dt.argtypes = [C.c_void_p]
        _lib.refshim_timer.restype = C.c_double; _lib.refshim_timer.argtypes = [C.c_void_p, C.c_int]
        _lib.refshim_box.restype = None
        _lib.refshim_box.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.refshim_getd.restype = C.c_int
        _lib.refshim_getd.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p]
        _lib.refshim_geti.restype = C.c_int
        _lib.refshim_geti.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p]
        _lib.refshim_cutneigh.restype = None
        _lib.refshim_cutneigh.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        _lib.refshim_has_full.restype = C.c_int; _lib.refshim_has_full.argtypes = [C.c_void_p]
        _lib.refshim_virial.restype = None; _lib.refshim_virial.argtypes = [C.c_void_p, C.c_void_p]
        _lib.refshim_neigh_full.restype = C.c_longlong
        _lib.refshim_neigh_full.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]
    return _lib


D3 = ("x", "v", "f", "vest", "colorgradient")
D1 = ("rho", "drho", "e", "de", "cv", "rmass")
I1 = ("type", "mask", "tag")


class RefLammps:
    def __init__(self, quiet=True):
        self.lib = _load()
        self.p = self.lib.refshim_open(1 if quiet else 0)

    def close(self):
        if self.p:
            self.lib.refshim_close(self.p); self.p = None

    def command(self, cmd):
        for line in cmd.strip().splitlines():
            line = line.strip()
            if line and not line.startswith("#"):
                self.lib.refshim_command(self.p, line.encode())

    def file(self, fn):
        self.lib.refshim_file(self.p, fn.encode())

    nlocal = property(lambda s: s.lib.refshim_nlocal(s.p))
    nghost = property(lambda s: s.lib.refshim_nghost(s.p))
    ntypes = property(lambda s: s.lib.refshim_ntypes(s.p))
    ntimestep = property(lambda s: s.lib.refshim_ntimestep(s.p))
    nbuilds = property(lambda s: s.lib.refshim_nbuilds(s.p))
    ndanger = property(lambda s: s.lib.refshim_ndanger(s.p))
    dt = property(lambda s: s.lib.refshim_dt(s.p))

    def loop_time(self):
        return self.lib.refshim_timer(self.p, 0)   # TIME_LOOP

    def box(self):
        lo = np.zeros(3); hi = np.zeros(3); per = np.zeros(3, np.int32); dim = C.c_int()
        self.lib.refshim_box(self.p, lo.ctypes.data, hi.ctypes.data, per.ctypes.data, C.byref(dim))
        return lo, hi, per, dim.value

    def get(self, name, ghost=False, multiphase=True):
        n = self.nlocal + (self.nghost if ghost else 0)
        if name == "mass":
            out = np.zeros(self.ntypes + 1)
            self.lib.refshim_getd(self.p, b"mass", 0, out.ctypes.data); return out
        if name in I1:
            out = np.zeros(n, np.int32)
            rc = self.lib.refshim_geti(self.p, name.encode(), int(ghost), out.ctypes.data)
        else:
            out = np.zeros((n, 3) if name in D3 else n)
            rc = self.lib.refshim_getd(self.p, name.encode(), int(ghost), out.ctypes.data)
        if rc < 0:
            raise KeyError(name)
        return out

    def state(self, ghost=False, multiphase=True):
        names = ["x", "v", "f", "vest", "rho", "drho", "e", "de", "cv", "type", "mask", "tag"]
        if multiphase:
            names += ["colorgradient", "rmass"]
        return {k: self.get(k, ghost) for k in names}

    def virial(self):
        """force->pair->virial (xx yy zz xy xz yz) of the last thermo step"""
        v = np.zeros(6)
        self.lib.refshim_virial(self.p, v.ctypes.data_as(C.c_void_p))
        return v

    def cutneigh(self):
        n1 = self.ntypes + 1
        cn = np.zeros((n1, n1)); cmax = C.c_double(); skin = C.c_double(); cg = C.c_double()
        ev = C.c_int(); de = C.c_int(); ch = C.c_int()
        self.lib.refshim_cutneigh(self.p, cn.ctypes.data, C.byref(cmax), C.byref(skin), C.byref(ev), C.byref(de),
                                  C.byref(ch), C.byref(cg))
        return dict(cutneighsq=cn, cutneighmax=cmax.value, skin=skin.value, every=ev.value, delay=de.value,
                    check=ch.value, cutghost=cg.value)

    def neighbor_list(self):
        """full list of the last build -> (numneigh, jtag, jimage), rows sorted by (tag,image),
        image = (px+1)+3(py+1)+9(pz+1) from the ghost's offset to its owner, 13 for owned"""
        nl = self.nlocal
        num = np.zeros(nl, np.int32)
        tot = self.lib.refshim_neigh_full(self.p, num.ctypes.data, 0, None)
        if tot < 0:
            raise RuntimeError("no built full list in the reference")
        j = np.zeros(max(tot, 1), np.int32)
        self.lib.refshim_neigh_full(self.p, num.ctypes.data, tot, j.ctypes.data)
        j = j[:tot]
        x = self.get("x", True); tag = self.get("tag", True)
        lo, hi, per, dim = self.box(); prd = hi - lo
        owner = np.zeros(tag.max() + 1, np.int64); owner[tag[:nl]] = np.arange(nl)
        p = np.rint((x - x[owner[tag]]) / prd).astype(np.int64)          # (nall,3) in {-1,0,1}
        img = (p[:, 0] + 1) + 3 * (p[:, 1] + 1) + 9 * (p[:, 2] + 1)
        jt = tag[j].astype(np.int64); ji = img[j]
        rows = np.repeat(np.arange(nl), num)
        if not self.lib.refshim_has_full(self.p):
            # only a half list was built (no full-list style in the deck): every pair appears once
            # (neigh_half_bin.cpp half_bin_newton); mirror it to recover the full list
            mrows = owner[tag[j]]; mt = tag[rows].astype(np.int64); mi = 26 - ji
            rows = np.concatenate([rows, mrows]); jt = np.concatenate([jt, mt]); ji = np.concatenate([ji, mi])
            num = np.bincount(rows, minlength=nl).astype(np.int32)
        order = np.lexsort((ji, jt, rows))
        return num, jt[order].astype(np.int32), ji[order].astype(np.int32)
