"""Contact workload (BASELINE config I8: hex projectile into hex plate, penalty contact) — reported separately from
the headline metric because its work is data dependent.

  python scripts/bench_contact.py [--plate 400,400,48] [--proj 68,68,68] [--steps 60] [--mu 0.0]
Prints one JSON line: ms/step per kernel class (CUDA events around every launch), hits and candidate tests per step."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--plate", default="400,400,48")
    ap.add_argument("--proj", default="68,68,68")
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=12)
    ap.add_argument("--mu", type=float, default=0.0)
    args = ap.parse_args()
    from hakai_fem_b200.engine import Engine
    from hakai_fem_b200.mesh import ImpactDeck
    from hakai_fem_b200.model_setup import prepare, configure_engine
    t0 = time.perf_counter()
    deck = ImpactDeck(plate=tuple(int(v) for v in args.plate.split(",")), proj=tuple(int(v) for v in args.proj.split(",")))
    model = deck.build_model()
    st = prepare(model)
    t_setup = time.perf_counter() - t0
    t0 = time.perf_counter()
    g = configure_engine(Engine, st, contact_myu=args.mu)
    t_engine = time.perf_counter() - t0
    g.step(1, args.warmup)                       # gap closes after 10 steps
    c0 = g.counters()
    g.profile(True)
    t0 = time.perf_counter()
    nd = g.step(args.warmup + 1, args.steps)
    wall = time.perf_counter() - t0
    ms, n = g.profile_read()
    c1 = g.counters()
    F = g.download_ex(fields=("external_force",))["external_force"]
    out = {
        "workload": f"impact plate {args.plate} + projectile {args.proj}, mu={args.mu}", "elements": int(model.nElement),
        "nodes": int(model.nNode), "pairs": [dict(nn_i=len(c.c_nodes_i), nn_j=len(c.c_nodes_j), nTri=len(c.c_triangles)) for c in st.CT],
        "steps": args.steps, "wall_ms_per_step": wall / args.steps * 1e3,
        "contact_ms_per_step": ms[0] / max(n[0], 1), "nodal_ms_per_step": ms[1] / max(n[1], 1),
        "element_ms_per_step": ms[2] / max(n[2], 1),
        "element_steps_per_s": model.nElement * args.steps / wall,
        "hits_per_step": float(c1[1] - c0[1]) / args.steps, "candidate_tests_per_step": float(c1[2] - c0[2]) / args.steps,
        "deleted": int(nd), "fixed_point_overflows": int(c1[5]), "max_contact_force": float(np.abs(F).max()),
        "host_setup_s": t_setup, "engine_setup_s": t_engine,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
