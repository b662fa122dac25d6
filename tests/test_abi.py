"""CPU: the C-ABI library loads and exports every symbol include/b200_sph.h declares;
no compute call is made (there is no GPU here and no CPU fallback in the product)."""
import ctypes
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = importlib.import_module("lammps-sph-multiphase_b200")


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "b200_sph.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_header():
    import __graft_entry__
    __graft_entry__.build()
    lib = ctypes.CDLL(pkg.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    # the ctypes view covers the header one to one
    assert sorted("b200_" + n for n in pkg._abi.ABI_SYMBOLS) == syms


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    api = pkg.load()
    h = ctypes.c_void_p()
    assert api.create(ctypes.byref(h), 0) < 0
    assert b"no CUDA device" in api.last_error()


def test_oracle_exports_same_abi():
    import harness
    api = harness.oracle_api()
    assert api.version().startswith(b"sph_oracle")
