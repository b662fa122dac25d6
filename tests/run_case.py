"""debug helper: python tests/run_case.py <case> [nsteps]  -- runs one parity case on the GPU engine"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib, cases, harness
pkg = importlib.import_module("lammps-sph-multiphase_b200")
name = sys.argv[1]
case = cases.CASES[name]; g = harness.load_golden(name)
sim = pkg.B200Sim(case.deck())
sim.set_atoms(**harness.state_from(g, "init_", case.multiphase))
sim.setup(); print("setup ok", sim.natoms(), sim.counters())
n = int(sys.argv[2]) if len(sys.argv) > 2 else case.nsteps
sim.run(n); a = sim.get_atoms(); print("run ok", sim.counters())
