"""Driver of the C-ABI (include/b200_sph.h): pushes a parsed Deck and the
per-atom arrays across the boundary and exposes Verlet::setup / run and the
stage-level hooks.  This is the Python twin of the LAMMPS `/b200` shells in
lammps/USER-B200: it holds no physics and no CPU fallback -- every call goes
to the shared library it was given (the product: libb200sph.so).
"""
import ctypes as C
import numpy as np
from . import _abi
from ._abi import Atoms, PairDesc, PhaseChangeDesc, c_double_p, c_int_p


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_int_p)


class Sim:
    """One engine instance (one GPU / one MPI rank)."""

    def __init__(self, api, deck, device=0, brick=None, nccl_id=None):
        """brick / nccl_id: this rank's place in a multi-GPU decomposition (parallel.Brick, parallel.nccl_id)"""
        self.api, self.deck, self.brick = api, deck, brick
        h = C.c_void_p()
        api.check(api.create(C.byref(h), device))
        self.h = h
        self._keep = []
        if brick is not None and brick.world > 1:
            grid = np.array(brick.grid, np.int32); loc = np.array(brick.myloc, np.int32); nb = np.array(brick.procneigh, np.int32)
            api.check(api.comm_init(h, brick.world, brick.rank, _ip(grid), _ip(loc), _ip(nb), nccl_id))
        self._configure()

    def request_virial(self):
        """arm Pair::virial_fdotr_compute for the force evaluation that ends the next setup() / run()"""
        self.api.check(self.api.request_virial(self.h))

    def virial(self):
        """pair virial xx yy zz xy xz yz of that evaluation (this rank's share)"""
        v = np.zeros(6)
        self.api.check(self.api.get_virial(self.h, _dp(v)))
        return v

    def box(self):
        """domain->boxlo / boxhi as the engine holds them (move under boundary s / m)"""
        lo, hi = np.zeros(3), np.zeros(3)
        self.api.check(self.api.get_box(self.h, _dp(lo), _dp(hi)))
        return lo, hi

    def timestep(self):
        """update->dt as the engine holds it (changes under fix dt/reset)"""
        dt = C.c_double()
        self.api.check(self.api.get_timestep(self.h, C.byref(dt)))
        return dt.value

    def set_time(self, atime, atimestep, laststep):
        """Update::atime / atimestep and FixDtReset::laststep as the caller holds them (before setup)"""
        self.api.check(self.api.set_time(self.h, float(atime), int(atimestep), int(laststep)))

    def time(self):
        """(atime, atimestep, laststep): elapsed-time bookkeeping under fix dt/reset (update.cpp:480-484, fix_dt_reset.cpp:175-181);
        thermo's `time` is atime + (ntimestep - atimestep) * dt"""
        a, s, l = C.c_double(), C.c_longlong(), C.c_longlong()
        self.api.check(self.api.get_time(self.h, C.byref(a), C.byref(s), C.byref(l)))
        return a.value, s.value, l.value

    def close(self):
        if self.h:
            self.api.destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------
    def _configure(self):
        api, d, h = self.api, self.deck, self.h
        if not d._initd:
            d.init()
        ck = api.check
        lo = np.array(d.boxlo, np.float64); hi = np.array(d.boxhi, np.float64)
        per = np.array(d.periodicity, np.int32)
        slo = np.array(self.brick.sublo, np.float64) if self.brick is not None else lo
        shi = np.array(self.brick.subhi, np.float64) if self.brick is not None else hi
        ck(api.domain(h, d.dimension, _dp(lo), _dp(hi), _ip(per), _dp(slo), _dp(shi)))
        if d.shrink:       # boundary s / m: the engine re-fits the box on every rebuild (Domain::reset_box)
            bnd = np.array(d.boundary, np.int32).reshape(6); small = np.array(d.small, np.float64)
            mb = np.array(d.minbox, np.float64).reshape(6)
            ck(api.boundary(h, _ip(bnd), _dp(small), _dp(mb)))
        mass = np.ascontiguousarray(d.mass_, np.float64)
        ck(api.atom_style(h, int(d.multiphase), d.ntypes, _dp(mass)))
        cn = np.ascontiguousarray(d.cutneighsq, np.float64)
        ck(api.neighbor(h, d.skin, d.every, d.delay, d.check, _dp(cn), d.cutneighmax, d.cutghost))
        ck(api.timestep(h, d.dt, d.ftm2v, d.ntimestep))
        ck(api.comm_modify(h, d.ghost_velocity))
        ck(api.atom_modify(h, d.sortfreq, d.sort_binsize))
        ck(api.pair_clear(h))
        for s in d.styles:
            t = {k: np.ascontiguousarray(getattr(s, k)) for k in
                 ("mapped", "cut", "cutsq", "rho0", "B", "soundspeed", "gamma", "rbackground",
                  "viscosity", "alpha", "tc", "fixflag")}
            self._keep.append(t)
            pd = PairDesc(s.style, s.nstep, _ip(t["mapped"]), _dp(t["cut"]), _dp(t["cutsq"]), _dp(t["rho0"]),
                          _dp(t["B"]), _dp(t["soundspeed"]), _dp(t["gamma"]), _dp(t["rbackground"]),
                          _dp(t["viscosity"]), _dp(t["alpha"]), _dp(t["tc"]), _ip(t["fixflag"]))
            ck(api.pair_add(h, C.byref(pd)))
        ck(api.fix_clear(h))
        for style, bit, arg in d.fixes:
            if style == "meso":
                ck(api.fix_meso(h, bit))
            elif style == "meso/stationary":
                ck(api.fix_meso_stationary(h, bit))
            elif style == "gravity":
                ck(api.fix_gravity(h, bit, *arg))
            elif style == "enforce2d":
                ck(api.fix_enforce2d(h, bit))
            elif style == "dt/reset":
                ck(api.fix_dt_reset(h, bit, *arg))
            elif style == "setmesode":
                value, kind, reg = arg
                r = np.array(reg, np.float64)
                ck(api.fix_setmesode(h, bit, value, kind, _dp(r)))
            elif style == "setforce":
                sets = np.array(arg[0], np.int32); vals = np.array(arg[1], np.float64)
                ck(api.fix_setforce(h, bit, _ip(sets), _dp(vals)))
            elif style == "setmeso":
                which, value, kind, reg, inside = arg
                r = np.array(reg, np.float64)
                ck(api.fix_setmeso(h, bit, which, value, kind, _dp(r), inside))
            elif style == "setmeso/var":
                which, formula, kind, reg, inside = arg
                r = np.array(reg, np.float64)
                ck(api.fix_setmeso_var(h, bit, which, formula.encode(), kind, _dp(r), inside))
            elif style == "addforce":
                vals = np.array(arg[0], np.float64)
                forms = (C.c_char_p * 3)(*[f.encode() if f is not None else None for f in arg[1]])
                ck(api.fix_addforce(h, bit, _dp(vals), forms))
            elif style == "phase_change":
                pc = PhaseChangeDesc(groupbit=bit, **arg)
                ck(api.fix_phase_change(h, C.byref(pc)))

    # ------------------------------------------------------------------
    @staticmethod
    def _bundle(n, fields, out=False):
        a, keep = Atoms(), {}
        for k in _abi.ATOM_FIELDS_D3 + _abi.ATOM_FIELDS_D1 + _abi.ATOM_FIELDS_I:
            v = fields.get(k)
            if v is None:
                continue
            if k in _abi.ATOM_FIELDS_I:
                arr = np.ascontiguousarray(v, np.int32)
                assert arr.shape == (n,), (k, arr.shape)
                setattr(a, k, _ip(arr))
            else:
                arr = np.ascontiguousarray(v, np.float64)
                assert arr.shape == ((n, 3) if k in _abi.ATOM_FIELDS_D3 else (n,)), (k, arr.shape)
                setattr(a, k, _dp(arr))
            keep[k] = arr
        return a, keep

    def set_atoms(self, **fields):
        """fields: x (n,3), v, vest, rho, e, cv, rmass, colorgradient, type, mask, tag"""
        n = len(fields["x"])
        a, keep = self._bundle(n, fields)
        self.api.check(self.api.set_atoms(self.h, n, C.byref(a)))

    def natoms(self):
        nl, ng = C.c_int(), C.c_int()
        self.api.check(self.api.get_natoms(self.h, C.byref(nl), C.byref(ng)))
        return nl.value, ng.value

    def get_atoms(self, names=("x", "v", "vest", "f", "rho", "drho", "e", "de", "cv", "rmass", "colorgradient",
                               "type", "mask", "tag"), out=None):
        """out: caller-owned arrays (name -> ndarray of at least nlocal rows, e.g. pinned host memory, as the LAMMPS
        shell hands the engine its own atom arrays); views of their first nlocal rows are filled and returned"""
        n, _ = self.natoms()
        if out is not None:
            views = {k: out[k][:n] for k in names}
            a, keep = self._bundle(n, views)
            for k in names:
                assert keep[k] is views[k] or np.shares_memory(keep[k], views[k]), k + ": dtype / layout forces a copy"
            self.api.check(self.api.get_atoms(self.h, n, C.byref(a)))
            return views
        out = {}
        for k in names:
            if k in _abi.ATOM_FIELDS_I:
                out[k] = np.zeros(n, np.int32)
            elif k in _abi.ATOM_FIELDS_D3:
                out[k] = np.zeros((n, 3), np.float64)
            else:
                out[k] = np.zeros(n, np.float64)
        a, keep = self._bundle(n, out)
        self.api.check(self.api.get_atoms(self.h, n, C.byref(a)))
        return keep

    # ------------------------------------------------------------------
    def setup(self):
        self.api.check(self.api.setup(self.h))

    def run(self, n):
        self.api.check(self.api.run(self.h, int(n)))

    def sync(self):
        self.api.check(self.api.sync(self.h))

    def call(self, name, *args):
        return self.api.check(getattr(self.api, name)(self.h, *args))

    def neigh_decide(self):
        r = C.c_int()
        self.api.check(self.api.neigh_decide(self.h, C.byref(r)))
        return r.value

    def neighbor_list(self):
        """-> (numneigh[nlocal], jtag[], jimage[]) rows sorted by (tag,image), LAMMPS local order"""
        n, _ = self.natoms()
        num = np.zeros(n, np.int32)
        self.api.check(self.api.get_neighbor_list(self.h, n, _ip(num), 0, None, None))
        tot = int(num.sum())
        jt = np.zeros(max(tot, 1), np.int32); ji = np.zeros(max(tot, 1), np.int32)
        self.api.check(self.api.get_neighbor_list(self.h, n, _ip(num), tot, _ip(jt), _ip(ji)))
        return num, jt[:tot], ji[:tot]

    def counters(self):
        c = (C.c_longlong * 8)()
        self.api.check(self.api.get_counters(self.h, c))
        names = ("launches", "builds", "steps", "max_neighbors", "nghost", "row_stride", "inserted", "dangerous")
        return dict(zip(names, [int(v) for v in c]))

    def set_timing(self, on=True):
        self.api.check(self.api.set_timing(self.h, int(on)))

    def timers(self, n=32):
        ms = (C.c_double * n)(); calls = (C.c_longlong * n)()
        self.api.check(self.api.get_timers(self.h, n, ms, calls))
        out = {}
        for i in range(n):
            nm = self.api.timer_name(i)
            if nm:
                out[nm.decode()] = (ms[i], int(calls[i]))
        return out
