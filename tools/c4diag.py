import importlib, sys, os
import numpy as np
ROOT='/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+'/tests')
import cases, harness
from util import relerr
import test_gpu_fullsize as T
pkg = importlib.import_module("lammps-sph-multiphase_b200")
nx=int(sys.argv[1]) if len(sys.argv)>1 else 100
atoms=T._mp_atoms(nx,"c4"); n0=len(atoms["type"])
res={}
for nsteps in (1,2,3):
    out=[]
    for mk in (pkg.B200Sim, harness.oracle_sim):
        sim=mk(cases._bubble("c4",3,nx,nsteps).deck()); sim.set_atoms(**atoms); sim.setup(); sim.run(nsteps)
        out.append((sim.get_atoms(), sim.counters())); sim.close()
    (a,ca),(b,cb)=out
    xa,xb=a["x"][n0:],b["x"][n0:]
    print("steps",nsteps,"inserted",ca["inserted"],cb["inserted"], flush=True)
    from scipy.spatial import cKDTree
    if len(xa) and len(xb):
        d,_=cKDTree(xb).query(xa); print("  engine insertions without oracle partner within 1e-9:", int((d>1e-9).sum()), "max d of matched", d[d<=1e-9].max() if (d<=1e-9).any() else None)
        d2,_=cKDTree(xa).query(xb); print("  oracle insertions without engine partner:", int((d2>1e-9).sum()))
        m=min(len(xa),len(xb)); same=np.abs(xa[:m]-xb[:m]).max(1)<1e-9
        first=np.argmin(same) if not same.all() else -1
        print("  first differing insertion index (in order):", first, "of", m)
    na=min(len(a["type"]),len(b["type"]))
    print("  owned fields first n0:", {k:"%.1e"%relerr(a[k][:n0],b[k][:n0]) for k in ("x","v","f","rho","e","de","rmass")})
    print("  types differing among first n0:", int((a["type"][:n0]!=b["type"][:n0]).sum()))
