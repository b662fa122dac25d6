"""Build libb200sph.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = ["b200_sph.cu"]
HDR = ["b200_common.cuh", "b200_neigh.cuh", "b200_pair.cuh", "b200_tile.cuh", "b200_fix.cuh", "b200_phase.cuh", "b200_comm.cuh", "b200_lj.cuh", "b200_expr.cuh", "../../include/b200_sph.h"]
OUT = os.path.join(HERE, "libb200sph.so")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
         "-shared", "-cudart", "static", "-ldl"]


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SRC + HDR + ["build.py"])


def build(verbose=False, force=False):
    if not force and not stale():
        return OUT
    cmd = ["nvcc"] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + [os.path.join(HERE, f) for f in SRC]
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(verbose="-v" in sys.argv, force=True)
