"""The C-ABI library loads and exports every symbol include/hakai_b200.h declares; without a CUDA device
hk_create fails loudly (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hakai_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hk_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ("hk_create", "hk_set_mesh", "hk_add_material", "hk_add_bc", "hk_add_ic", "hk_add_contact_pair",
              "hk_finalize", "hk_step", "hk_download", "hk_destroy", "hk_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from hakai_fem_b200.engine import load_library, EXPORTS
    lib = load_library()
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/hakai_b200.h but not exported"
    assert sorted("hk_" + e for e in EXPORTS) == declared_symbols()


def test_oracle_exports_same_abi(oracle_lib):
    for s in declared_symbols():
        assert hasattr(oracle_lib, "hko_" + s[3:])


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from hakai_fem_b200.engine import Engine, HakaiError
    with pytest.raises(HakaiError) as ei:
        Engine(d_time=1e-7)
    assert "no CUDA device" in str(ei.value) or "fallback" in str(ei.value)


def test_params_struct_size_matches():
    from hakai_fem_b200.engine import load_library, HkParams
    lib = load_library()
    p = HkParams()
    assert lib.hk_default_params(C.byref(p)) == 0
    assert p.struct_size == C.sizeof(HkParams)
    assert p.contact_myu == 0.25 and p.contact_d_lim_factor == 0.3 and p.contact_ddiv_other == 1.1
