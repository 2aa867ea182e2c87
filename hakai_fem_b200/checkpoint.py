"""Checkpoint / resume of the engine state (the reference has none — its state is the local
variables of hakai(), J2:220-230, 430-465).

A checkpoint is the loop-carried state after step t, in the reference's own array layouts:
    disp, disp_pre, velo, Q (fn) · integ_stress, integ_strain (6,nip) · integ_eq_plastic_strain,
    integ_yield_stress (nip) · element_flag (nElement) · deleted ids in deletion order · t
Everything else is recomputed from these inside a step (external_force, d_disp, position, triax), so a run
resumed from a checkpoint continues bit-for-bit like the uninterrupted one (tests/test_checkpoint.py).
The contact surfaces grown by erosion are not stored: `hk_apply_deleted` replays add_surface_triangle
(J2:2167-2245) for the recorded ids in their original order on the freshly configured engine.
"""
from __future__ import annotations

import os

import numpy as np

_FIELDS = ("disp", "disp_pre", "velo", "Q", "integ_stress", "integ_strain", "integ_eq_plastic_strain",
           "integ_yield_stress", "element_flag")
FORMAT = 1


def save_checkpoint(engine, path: str, t: int) -> None:
    """Writes the state after step `t` (call between steps; the engine is synchronised by the downloads)."""
    d = engine.download(fields=("disp", "velo", "integ_stress", "integ_strain", "integ_eq_plastic_strain",
                                "element_flag"))
    x = engine.download_ex(fields=("disp_pre", "Q", "integ_yield_stress"))
    out = {k: (d[k] if k in d else x[k]) for k in _FIELDS}
    # written to the EXACT path given (np.savez on a file name would append ".npz") through a temporary file that is
    # renamed over it: a crash during the write leaves the previous restart point intact
    tmp = "%s.tmp.%d" % (path, os.getpid())
    with open(tmp, "wb") as f:
        np.savez(f, t=np.int64(t), format=np.int64(FORMAT), nNode=np.int64(engine.nNode),
                 nElement=np.int64(engine.nElement), deleted_ids=engine.deleted_ids(), **out)
        f.flush()
        os.fsync(f.fileno())
    os.replace(tmp, path)


def load_checkpoint(engine, path: str) -> int:
    """Restores a checkpoint into a freshly configured engine of the SAME deck; returns t (resume at t + 1)."""
    with np.load(path) as z:
        if int(z["format"]) != FORMAT:
            raise ValueError("unknown checkpoint format %d" % int(z["format"]))
        if int(z["nNode"]) != engine.nNode or int(z["nElement"]) != engine.nElement:
            raise ValueError("checkpoint belongs to another mesh (%d nodes, %d elements)"
                             % (int(z["nNode"]), int(z["nElement"])))
        if len(engine.deleted_ids()):
            raise ValueError("load_checkpoint needs a freshly configured engine")
        engine.upload_state(**{k: z[k] for k in _FIELDS})
        ids = z["deleted_ids"]
        if len(ids):
            engine.apply_deleted(ids)
        return int(z["t"])
