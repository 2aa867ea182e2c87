"""Checkpoint / resume (hakai_fem_b200/checkpoint.py): a resumed run continues bit-for-bit like the uninterrupted
one — fields, deleted ids in order, and the contact surfaces grown by erosion."""
import numpy as np
import pytest

from hakai_fem_b200.checkpoint import save_checkpoint, load_checkpoint
from hakai_fem_b200.model_setup import configure_engine
from oracle.oracle_engine import OracleEngine

from .emu.emu_engine import EmuEngine
from .parity_cases import check_resume, erosion_setup as _erosion_setup, fracture_setup as _fracture_setup


@pytest.mark.parametrize("engine_cls", [OracleEngine, EmuEngine], ids=["oracle", "kernels"])
def test_resume_contact_erosion_is_bitwise(engine_cls, tmp_path):
    a, b, n_before = check_resume(engine_cls, _erosion_setup, 57, 120, tmp_path)
    assert 0 < n_before < len(a.deleted_ids()), "checkpoint must fall between deletions"
    for c in range(2):
        pa, pb = a.contact_pair(c), b.contact_pair(c)
        for k in ("c_nodes_i", "c_nodes_j", "c_triangles", "c_triangles_eleid"):
            assert np.array_equal(pa[k], pb[k]), f"pair {c} {k}"


@pytest.mark.parametrize("engine_cls", [OracleEngine, EmuEngine], ids=["oracle", "kernels"])
def test_resume_fracture_block_is_bitwise(engine_cls, tmp_path):
    a, _, n_before = check_resume(engine_cls, _fracture_setup, 61, 100, tmp_path)
    assert 0 < n_before < len(a.deleted_ids())


def test_checkpoint_rejects_other_mesh_and_used_engine(tmp_path):
    a = configure_engine(OracleEngine, _fracture_setup())
    a.step(1, 200)
    assert len(a.deleted_ids()) > 0
    path = str(tmp_path / "ck.npz")
    save_checkpoint(a, path, 200)
    with pytest.raises(ValueError):
        load_checkpoint(configure_engine(OracleEngine, _erosion_setup()), path)
    with pytest.raises(ValueError):
        load_checkpoint(a, path)                                  # not fresh: it already has a deletion history


def test_checkpoint_is_written_to_the_exact_path_atomically(tmp_path):
    """`checkpoint='run.ckpt'` must create run.ckpt (np.savez on a name would write run.ckpt.npz and resume
    would not find it); the file is replaced in one rename, no temporary is left behind; and the host driver writes
    checkpoints even when it writes no frames."""
    import os
    a = configure_engine(OracleEngine, _fracture_setup())
    a.step(1, 10)
    path = str(tmp_path / "run.ckpt")
    save_checkpoint(a, path, 10)
    save_checkpoint(a, path, 10)                                   # overwrite in place
    assert sorted(os.listdir(tmp_path)) == ["run.ckpt"]
    b = configure_engine(OracleEngine, _fracture_setup())
    assert load_checkpoint(b, path) == 10
    assert np.array_equal(a.download()["disp"], b.download()["disp"])
    from hakai_fem_b200.host import hakai
    from hakai_fem_b200.mesh import StretchDeck
    deck_path = str(tmp_path / "d.inp")
    StretchDeck(3, 3, 4, jitter=0.05, n_steps=40).write_inp(deck_path)
    ck = str(tmp_path / "frames_off.ckpt")
    hakai(deck_path, str(tmp_path / "out"), engine_cls=OracleEngine, output_num=4, write_frames=False, verbose=False,
          checkpoint=ck, checkpoint_frames=1)
    assert os.path.exists(ck)
