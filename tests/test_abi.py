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


def test_package_install_script(tmp_path):
    """USER-B200/Install.sh in a mock LAMMPS src/: install twice (idempotent), then uninstall"""
    import shutil
    import subprocess
    pk = os.path.join(ROOT, "lammps-sph-multiphase_b200", "lammps", "USER-B200")
    src = tmp_path / "src"
    shutil.copytree(pk, src / "USER-B200")
    (src / "Makefile.package").write_text("PKG_INC = -DX \nPKG_PATH = \nPKG_LIB = -lz \n")
    (src / "Makefile.package.settings").write_text("# settings\n")
    run = lambda mode: subprocess.run(["sh", "Install.sh", str(mode)], cwd=src / "USER-B200", capture_output=True, text=True)
    assert run(1).returncode != 0 and not (src / "verlet_b200.cpp").exists()       # USER-SPH (multiphase) must be there first
    (src / "pair_sph_taitwater_multiphase.cpp").write_text("")
    for _ in range(2):
        assert run(1).returncode == 0
    shells = sorted(f for f in os.listdir(pk) if f.endswith((".h", ".cpp")))
    assert all((src / f).exists() for f in shells) and len(shells) == 7
    mk = (src / "Makefile.package").read_text()
    assert mk.count("$(b200sph_INC)") == 1 and mk.count("$(b200sph_LIB)") == 1 and "-DX" in mk and "-lz" in mk
    assert (src / "Makefile.package.settings").read_text().count("Makefile.b200sph") == 1
    assert run(0).returncode == 0
    assert not any((src / f).exists() for f in shells)
    assert "b200sph" not in (src / "Makefile.package").read_text() + (src / "Makefile.package.settings").read_text()
    assert "-DX" in (src / "Makefile.package").read_text()
