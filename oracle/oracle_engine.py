"""ctypes wrapper of the CPU oracle (oracle/libhakai_oracle.so, prefix hko_).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  It reuses the product's generic ctypes wrapper class so that the parity tests
drive oracle and CUDA engine through identical calls; the product never imports this module.
"""
import ctypes as C
import os
import subprocess

from hakai_fem_b200.engine import EngineBase

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhakai_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "hakai_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
    return _lib


class OracleEngine(EngineBase):
    builds_contact = False
    """CPU restatement of the reference step (HAKAI_j.jl:487-951)."""

    def __init__(self, **params):
        super().__init__(load(), "hko_", **params)
