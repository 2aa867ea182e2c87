"""Kernel-body parity on the CPU: the CUDA engine's per-thread kernel bodies and host plumbing, compiled
for the host (tests/emu, -DHK_EMU), against the oracle.  This validates logic in the GPU-less container;
the same cases run against the real CUDA build in test_gpu_parity.py (-m gpu)."""
import pytest

from . import parity_cases as pc
from .emu.emu_engine import EmuEngine


def test_roundtrip():
    pc.case_roundtrip(EmuEngine)


@pytest.mark.parametrize("ductile", [False, True])
def test_single_step_random_state(ductile):
    pc.case_single_step_random_state(EmuEngine, ductile)


def test_t5_first_1000_steps():
    pc.case_t5(EmuEngine, n_total=1000)


def test_fracture_block():
    pc.case_fracture_block(EmuEngine)


@pytest.mark.parametrize("mu", [0.0, 0.25])
def test_contact(mu):
    pc.case_contact(EmuEngine, mu)


def test_contact_penetration_clamp():
    pc.case_contact_clamp(EmuEngine)


def test_contact_single_step_exact():
    pc.case_contact_single_step(EmuEngine)


def test_contact_pair_surfaces(tmp_path):
    pc.case_contact_pair_surfaces(EmuEngine, tmp_path)


@pytest.mark.parametrize("name,n_steps,min_deleted", [("bullet_impact", 2500, 5), ("charpy", 300, 0), ("crash_tube", 200, 0),
                                                     ("metal_cutting", 1500, 1)])
def test_contact_built_on_device(name, n_steps, min_deleted):
    assert pc.case_build_contact(EmuEngine, name, n_steps) >= min_deleted


def test_bc_edge_cases():
    pc.case_bc_edge_cases(EmuEngine)


def test_state_summary():
    pc.case_state_summary(EmuEngine)


def test_node_output():
    pc.case_node_output(EmuEngine)


def test_contact_erosion():
    pc.case_contact_erosion(EmuEngine)


def test_exact_mode_bitwise():
    pc.case_exact_mode_bitwise(EmuEngine)
