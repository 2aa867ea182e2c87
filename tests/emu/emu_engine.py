"""Host-compiled debugging build of the CUDA kernel bodies (libhakai_emu.so, prefix hke_).

TESTS ONLY: lets the kernel logic and the engine plumbing be exercised in the GPU-less build container.
Never imported by the package; the product path is hakai_fem_b200.engine.Engine (CUDA, no fallback).
"""
import ctypes as C
import os
import subprocess

from hakai_fem_b200.engine import EngineBase

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhakai_emu.so")
_CSRC = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "hakai_fem_b200", "csrc")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _CSRC, "emu"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
    return _lib


class EmuEngine(EngineBase):
    def __init__(self, **params):
        super().__init__(load(), "hke_", **params)
