"""ctypes view of the C-ABI declared in include/b200_sph.h.

`bind(lib, prefix)` attaches prototypes to every exported entry point.  The
product only ever binds libb200sph.so with prefix "b200_"; the test-suite binds
the CPU oracle (same signatures, prefix "osph_") through the same function so
both sides are driven by identical calls.
"""
import ctypes as C

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_ll_p = C.POINTER(C.c_longlong)


class PairDesc(C.Structure):
    """b200_pair_desc"""
    _fields_ = [("style", C.c_int), ("nstep", C.c_int),
                ("mapped", c_int_p), ("cut", c_double_p), ("cutsq", c_double_p),
                ("rho0", c_double_p), ("B", c_double_p), ("soundspeed", c_double_p),
                ("gamma", c_double_p), ("rbackground", c_double_p),
                ("viscosity", c_double_p), ("alpha", c_double_p), ("tc", c_double_p),
                ("fixflag", c_int_p)]


class PhaseChangeDesc(C.Structure):
    """b200_phase_change_desc"""
    _fields_ = [("groupbit", C.c_int),
                ("Tc", C.c_double), ("Tt", C.c_double), ("Hwv", C.c_double), ("dr", C.c_double),
                ("to_mass", C.c_double), ("cutoff", C.c_double),
                ("from_type", C.c_int), ("to_type", C.c_int), ("nfreq", C.c_int), ("seed", C.c_int),
                ("energy_chance_flag", C.c_int),
                ("change_chance", C.c_double), ("phase_change_rate", C.c_double),
                ("maxattempt", C.c_int), ("first_step", C.c_longlong)]


class Atoms(C.Structure):
    """b200_atoms"""
    _fields_ = [("x", c_double_p), ("v", c_double_p), ("vest", c_double_p), ("f", c_double_p),
                ("rho", c_double_p), ("drho", c_double_p), ("e", c_double_p), ("de", c_double_p),
                ("cv", c_double_p), ("rmass", c_double_p), ("colorgradient", c_double_p),
                ("type", c_int_p), ("mask", c_int_p), ("tag", c_int_p)]


ATOM_FIELDS_D3 = ("x", "v", "vest", "f", "colorgradient")
ATOM_FIELDS_D1 = ("rho", "drho", "e", "de", "cv", "rmass")
ATOM_FIELDS_I = ("type", "mask", "tag")

_H = C.c_void_p
_PROTOS = {
    "create": (C.c_int, [C.POINTER(_H), C.c_int]),
    "destroy": (C.c_int, [_H]),
    "last_error": (C.c_char_p, []),
    "version": (C.c_char_p, []),
    "comm_unique_id": (C.c_int, [C.c_char_p]),
    "comm_init": (C.c_int, [_H, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p, C.c_char_p]),
    "domain": (C.c_int, [_H, C.c_int, c_double_p, c_double_p, c_int_p, c_double_p, c_double_p]),
    "boundary": (C.c_int, [_H, c_int_p, c_double_p, c_double_p]),
    "get_box": (C.c_int, [_H, c_double_p, c_double_p]),
    "atom_style": (C.c_int, [_H, C.c_int, C.c_int, c_double_p]),
    "neighbor": (C.c_int, [_H, C.c_double, C.c_int, C.c_int, C.c_int, c_double_p, C.c_double, C.c_double]),
    "timestep": (C.c_int, [_H, C.c_double, C.c_double, C.c_longlong]),
    "comm_modify": (C.c_int, [_H, C.c_int]),
    "atom_modify": (C.c_int, [_H, C.c_int, C.c_double]),
    "pair_clear": (C.c_int, [_H]),
    "pair_add": (C.c_int, [_H, C.POINTER(PairDesc)]),
    "fix_clear": (C.c_int, [_H]),
    "fix_meso": (C.c_int, [_H, C.c_int]),
    "fix_meso_stationary": (C.c_int, [_H, C.c_int]),
    "fix_gravity": (C.c_int, [_H, C.c_int, C.c_double, C.c_double, C.c_double]),
    "fix_phase_change": (C.c_int, [_H, C.POINTER(PhaseChangeDesc)]),
    "fix_setmeso": (C.c_int, [_H, C.c_int, C.c_int, C.c_double, C.c_int, c_double_p, C.c_int]),
    "fix_enforce2d": (C.c_int, [_H, C.c_int]),
    "fix_setforce": (C.c_int, [_H, C.c_int, c_int_p, c_double_p]),
    "fix_setmesode": (C.c_int, [_H, C.c_int, C.c_double, C.c_int, c_double_p]),
    "fix_setmeso_var": (C.c_int, [_H, C.c_int, C.c_int, C.c_char_p, C.c_int, c_double_p, C.c_int]),
    "fix_addforce": (C.c_int, [_H, C.c_int, c_double_p, C.POINTER(C.c_char_p)]),
    "formula_check": (C.c_int, [C.c_char_p, c_double_p, C.c_int, C.c_int, C.c_double, C.c_double, c_double_p]),
    "fix_dt_reset": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double]),
    "get_timestep": (C.c_int, [_H, C.POINTER(C.c_double)]),
    "set_time": (C.c_int, [_H, C.c_double, C.c_longlong, C.c_longlong]),
    "get_time": (C.c_int, [_H, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "request_virial": (C.c_int, [_H]),
    "get_virial": (C.c_int, [_H, c_double_p]),
    "set_atoms": (C.c_int, [_H, C.c_int, C.POINTER(Atoms)]),
    "get_natoms": (C.c_int, [_H, c_int_p, c_int_p]),
    "get_atoms": (C.c_int, [_H, C.c_int, C.POINTER(Atoms)]),
    "setup": (C.c_int, [_H]),
    "run": (C.c_int, [_H, C.c_int]),
    "initial_integrate": (C.c_int, [_H]),
    "final_integrate": (C.c_int, [_H]),
    "neigh_decide": (C.c_int, [_H, c_int_p]),
    "forward_comm": (C.c_int, [_H]),
    "reneighbor": (C.c_int, [_H]),
    "force_clear": (C.c_int, [_H]),
    "pair_compute": (C.c_int, [_H, C.c_int]),
    "pair_compute_all": (C.c_int, [_H]),
    "reverse_comm": (C.c_int, [_H]),
    "post_force": (C.c_int, [_H]),
    "get_neighbor_list": (C.c_int, [_H, C.c_int, c_int_p, C.c_longlong, c_int_p, c_int_p]),
    "get_counters": (C.c_int, [_H, c_ll_p]),
    "set_timing": (C.c_int, [_H, C.c_int]),
    "get_timers": (C.c_int, [_H, C.c_int, c_double_p, c_ll_p]),
    "timer_name": (C.c_char_p, [C.c_int]),
    "sync": (C.c_int, [_H]),
}

ABI_SYMBOLS = tuple(_PROTOS)


class Api:
    """Bound entry points of one shared library (attribute per ABI function)."""

    def __init__(self, lib, prefix):
        self.lib, self.prefix = lib, prefix
        for name, (res, args) in _PROTOS.items():
            fn = getattr(lib, prefix + name)   # AttributeError if the symbol is missing
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)

    def check(self, rc):
        if rc < 0:
            raise RuntimeError("%s: %s" % (self.prefix.rstrip("_"), self.last_error().decode()))
        return rc


def bind(lib, prefix):
    return Api(lib, prefix)
