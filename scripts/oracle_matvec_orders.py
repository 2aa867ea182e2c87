"""Bounds what the UNPINNED third-party arithmetic of the reference can change.

HAKAI_j.jl evaluates `Bfinal * d_u`, `Dmat * d_e_vec` and `Bfinal' * final_stress` (J2:1204-1205, 1330) with
StaticArrays, whose version is not pinned by the reference tree: depending on it the products are left-to-right sums of
products or muladd (FMA) chains.  The oracle implements both (HKO_MATVEC=plain | muladd).  This script runs the
reference's own Tensile5e deck for all 20 000 steps and a jittered ductile block both ways and reports the deletion
steps and the largest field differences — the uncertainty band any bit-level parity claim against the real Julia run
would carry.   python scripts/oracle_matvec_orders.py > profiles/r2_oracle_matvec_orders.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hakai_fem_b200.model_setup import prepare, configure_engine      # noqa: E402
from oracle.oracle_engine import OracleEngine                          # noqa: E402
from tests import util                                                 # noqa: E402


def run(st, n_steps, mode, chunk):
    os.environ["HKO_MATVEC"] = mode
    o = configure_engine(OracleEngine, st)
    first = None
    for t in range(0, n_steps, chunk):
        nd = o.step(t + 1, min(chunk, n_steps - t))
        if nd and first is None:
            first = int(o.deleted_steps()[0])
    d = o.download()
    return dict(disp=d["disp"], eps=d["integ_eq_plastic_strain"], stress=np.asarray(d["integ_stress"]),
                deleted=o.deleted_ids().tolist(), steps=o.deleted_steps().tolist(), first=first)


def compare(name, st, n_steps, chunk):
    a, b = run(st, n_steps, "plain", chunk), run(st, n_steps, "muladd", chunk)
    rel = lambda x, y: float(np.abs(x - y).max() / max(np.abs(x).max(), 1e-300))
    return {"deck": name, "steps": n_steps, "deleted_plain": a["deleted"][:8], "deleted_muladd": b["deleted"][:8],
            "n_deleted": [len(a["deleted"]), len(b["deleted"])],
            "deleted_sets_equal": sorted(a["deleted"]) == sorted(b["deleted"]),
            "first_deletion_step": [a["first"], b["first"]],
            "deletion_steps_equal": a["steps"] == b["steps"],
            "max_rel_diff": {"disp": rel(a["disp"], b["disp"]), "eq_plastic_strain": rel(a["eps"], b["eps"]),
                             "stress": rel(a["stress"], b["stress"])}}


def main():
    out = [compare("Tensile5e.inp (reference deck, 5 elements)", prepare(util.t5_model()), 20000, 1000)]
    deck = util.distorted_block(nx=5, ny=4, nz=6, jitter=0.05, ductile=True, strain_per_step=4e-4)
    out.append(compare("jittered ductile block 5x4x6, 70 steps", prepare(deck.build_model()), 70, 1))
    print(json.dumps({"what": "oracle with left-to-right products vs muladd chains in the three StaticArrays products of "
                              "cal_stress_hexa", "results": out}, indent=1))


if __name__ == "__main__":
    main()
