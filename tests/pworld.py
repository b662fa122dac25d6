"""Test infrastructure: the CPU oracle as P ranks of a LAMMPS-style brick decomposition, emulated in one process
(oracle/sph_oracle.c "P ranks in one process": exchange / borders / forward / reverse in lock step between the ranks' arrays).
It gives the expected values of the multi-GPU path for decks whose result depends on the decomposition (moving multiphase
decks read stale ghost rho / colorgradient); decomposition-independent decks must reproduce the 1-rank reference
fixtures, which is what pins the emulated collectives (tests/test_world_cpu.py)."""
import ctypes as C
import importlib

import numpy as np

import harness

pkg = importlib.import_module("lammps-sph-multiphase_b200")


class OracleWorld:
    def __init__(self, deck, world, grid=None, splits=None):
        self.api = harness.oracle_api()
        self.deck, self.world = deck, world
        self.bricks = [pkg.parallel.Brick(world, r, deck.boxlo, deck.boxhi, deck.dimension, grid, splits) for r in range(world)]
        self.sims = [pkg.Sim(self.api, deck, brick=b) for b in self.bricks]
        lib = self.api.lib
        for name in ("osph_world_setup", "osph_world_run"):
            getattr(lib, name).restype = C.c_int
        lib.osph_world_setup.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        lib.osph_world_run.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int]
        self._ranks = (C.c_void_p * world)(*[s.h for s in self.sims])

    @property
    def grid(self):
        return self.bricks[0].grid

    def set_atoms(self, **st):
        for sim, b in zip(self.sims, self.bricks):
            mine = b.owns(st["x"])
            sim.set_atoms(**{k: np.ascontiguousarray(v[mine]) for k, v in st.items()})

    def setup(self):
        self.api.check(self.api.lib.osph_world_setup(self._ranks, self.world))

    def run(self, nsteps):
        self.api.check(self.api.lib.osph_world_run(self._ranks, self.world, int(nsteps)))

    def natoms(self):
        return [s.natoms() for s in self.sims]

    def builds(self):
        return [s.counters()["builds"] for s in self.sims]

    def get_atoms(self):
        """all ranks' owned atoms, sorted by tag"""
        parts = [s.get_atoms() for s in self.sims]
        out = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
        order = np.argsort(out["tag"], kind="stable")
        return {k: v[order] for k, v in out.items()}

    def close(self):
        for s in self.sims:
            s.close()
