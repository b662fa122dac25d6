"""B200-native SPH hot path behind the reference's USER-SPH style names.

The package is a thin host layer over libb200sph.so (csrc/, sm_100a CUDA):
  deck.py    mirror of the reference's pair_style/pair_coeff/neighbor/fix semantics
  engine.py  driver of the C-ABI in include/b200_sph.h
There is no CPU implementation in here: `load()` raises if the CUDA library
has not been built (python __graft_entry__.py build).
"""
import ctypes
import os

from . import _abi
from .deck import Deck, DeckError, PAIR_STYLES
from .engine import Sim
from . import parallel

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200_LIB: another build of the same sources (A/B measurements of compile-time variants, tools/gpu_*.sh); not a fallback
LIB_PATH = os.environ.get("B200_LIB") or os.path.join(_HERE, "csrc", "libb200sph.so")
_api = None


def load():
    """bind libb200sph.so (RTLD_GLOBAL not needed); loud failure if it is not built"""
    global _api
    if _api is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libb200sph.so not built: run `python __graft_entry__.py build` "
                               "(there is no CPU fallback for the SPH hot path)")
        _api = _abi.bind(ctypes.CDLL(LIB_PATH), "b200_")
    return _api


def B200Sim(deck, device=0, brick=None, nccl_id=None):
    """the product entry point: one engine instance on cuda:<device> (optionally one brick of a multi-GPU run)"""
    return Sim(load(), deck, device, brick, nccl_id)
