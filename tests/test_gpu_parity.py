"""Parity of the CUDA engine (libhakai_b200.so, through the C ABI) against the CPU oracle.  -m gpu."""
import numpy as np
import pytest

from hakai_fem_b200.engine import Engine
from hakai_fem_b200.model_setup import prepare, configure_engine

from . import parity_cases as pc
from . import util

pytestmark = pytest.mark.gpu


def test_roundtrip():
    pc.case_roundtrip(Engine)


@pytest.mark.parametrize("ductile", [False, True])
def test_single_step_random_state(ductile):
    pc.case_single_step_random_state(Engine, ductile)


def test_t5_full_run():
    pc.case_t5(Engine, n_total=20000)


def test_fracture_block():
    pc.case_fracture_block(Engine)


@pytest.mark.parametrize("mu", [0.0, 0.25])
def test_contact(mu):
    pc.case_contact(Engine, mu)


def test_contact_penetration_clamp():
    pc.case_contact_clamp(Engine)


def test_contact_single_step_exact():
    pc.case_contact_single_step(Engine)


def test_contact_pair_surfaces(tmp_path):
    pc.case_contact_pair_surfaces(Engine, tmp_path)


@pytest.mark.parametrize("name,n_steps,min_deleted", [("bullet_impact", 2500, 5), ("charpy", 300, 0), ("crash_tube", 200, 0),
                                                     ("metal_cutting", 1500, 1)])
def test_contact_built_on_device(name, n_steps, min_deleted):
    assert pc.case_build_contact(Engine, name, n_steps) >= min_deleted


def test_bc_edge_cases():
    pc.case_bc_edge_cases(Engine)


def test_state_summary():
    pc.case_state_summary(Engine)


def test_node_output():
    pc.case_node_output(Engine)


def test_checkpoint_resume_bitwise(tmp_path):
    pc.case_checkpoint_resume(Engine, tmp_path)


def test_contact_erosion():
    pc.case_contact_erosion(Engine)


def test_bitwise_reproducible():
    """Two runs of the same deck give bit-identical state (no unordered float atomics anywhere)."""
    st, prm = pc.small_impact(0.25, plate=(16, 16, 3), proj=(5, 5, 5))
    outs = []
    for _ in range(2):
        g = configure_engine(Engine, st, **prm)
        g.step(1, 80)
        outs.append(util.full_state(g))
        g.close()
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k]), k


@pytest.mark.parametrize("name,n_steps,expect_deleted", [("bullet_impact", 3000, 13), ("metal_cutting", 3000, 30),
                                                          ("charpy", 1500, 0), ("car_crash_n2k", 1500, 0)])
def test_reference_deck(name, n_steps, expect_deleted):
    """The reference's own example decks through the CUDA engine vs the oracle (see tests/test_reference_decks.py)."""
    from .test_reference_decks import run_deck
    run_deck(Engine, name, n_steps, expect_deleted)


def test_exact_mode_bitwise():
    """On the B200: element_mode=1 reproduces the CPU oracle bit for bit (IEEE FP64, -fmad=false)."""
    pc.case_exact_mode_bitwise(Engine)


def test_host_driver_with_device_built_contact_writes_the_oracle_frames(tmp_path):
    """hakai(deck) with the CUDA engine — contact tables built by hk_build_contact (the default there), exposed faces
    and frames' nodal averages on the device — against the same driver on the CPU oracle with host-built tables."""
    import os
    from hakai_fem_b200.host import hakai
    from hakai_fem_b200.mesh import ImpactDeck
    from oracle.oracle_engine import OracleEngine
    from .test_multi_gloo import _vtk_sections
    deck_path = str(tmp_path / "deck.inp")
    ImpactDeck(plate=(8, 8, 2), proj=(3, 3, 3), v0=-900.0, n_steps=120,
               plate_ductile=[[0.02, 0.0, 30.0], [0.015, 0.4, 30.0]]).write_inp(deck_path)
    g, got = hakai(deck_path, str(tmp_path / "gpu"), output_num=6, verbose=False)
    o, ref = hakai(deck_path, str(tmp_path / "cpu"), engine_cls=OracleEngine, output_num=6, verbose=False)
    assert len(o.deleted_ids()) > 0 and np.array_equal(o.deleted_ids(), g.deleted_ids())
    assert len(got) == len(ref) == 7
    for c in range(2):
        po, pg = o.contact_pair(c), g.contact_pair(c)
        for k in po:
            assert np.array_equal(po[k], pg[k]), (c, k)
    for fa, fb in zip(ref, got):
        a, b = _vtk_sections(fa), _vtk_sections(fb)
        assert list(a) == list(b)
        for k in a:
            assert a[k].shape == b[k].shape, (os.path.basename(fa), k)
            if k in ("CELLS", "CELL_TYPES", "POINTS"):
                assert np.array_equal(a[k], b[k]), (os.path.basename(fa), k)
            elif k != "TRIAX_STRESS":                            # a ratio that is rounding noise while the plate is at rest
                scale = max(np.abs(a[k]).max(), 1e-300)
                assert np.allclose(a[k], b[k], rtol=0, atol=1e-5 * scale), (os.path.basename(fa), k, scale)
