"""A/B of element-kernel variants inside ONE process on one box (boxes differ by several %, runs on one box by ~1 %):
the deck is built once, every configuration gets its own engine, configurations are visited round-robin `--rounds`
times, each visit = warm-up into the plastic regime + `--steps` timed steps with CUDA events around every launch.

  python scripts/ab_element.py --configs "20,1,;12,1,;20,0,;20,1,red" [--workload W16] [--steps 30] [--rounds 2]
Build the comparison kernels first: make -C hakai_fem_b200/csrc ab (the shipped library holds the default only).
A configuration is variant,rec_soa,experiment (HK_ELEMENT_VARIANT, HK_REC_SOA, HK_EXPERIMENT).  Prints one JSON line per visit and a summary."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="20,1,;12,1,")
    ap.add_argument("--workload", default="W16")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=24)
    ap.add_argument("--rounds", type=int, default=2)
    args = ap.parse_args()
    import bench
    from hakai_fem_b200.engine import Engine
    from hakai_fem_b200.model_setup import configure_engine
    deck, kind = bench.make_deck(args.workload)
    st = bench.prepare_setup(deck)
    nE = st.model.nElement
    cfgs = [c.split(",") for c in args.configs.split(";") if c]
    res = {}
    for rnd in range(args.rounds):
        for v, blocked, exp in cfgs:
            os.environ["HK_ELEMENT_VARIANT"] = v
            os.environ["HK_REC_SOA"] = blocked
            os.environ["HK_EXPERIMENT"] = exp
            g = configure_engine(Engine, st)
            g.step(1, args.warmup)
            s = g.state_summary()
            g.profile(True)
            g.step(args.warmup + 1, args.steps)
            ms, n = g.profile_read()
            g.close()
            el, nd = ms[2] / max(n[2], 1), ms[1] / max(n[1], 1)
            key = f"v{v} rec_soa={blocked} {exp}".strip()
            res.setdefault(key, []).append((el, nd))
            print(json.dumps({"config": key, "round": rnd, "element_ms": el, "nodal_ms": nd,
                              "frac": 1904 * nE / (el * 1e-3) / 1e9 / 6448.4,
                              "plastic": s["yielded_points"] == 8 * s["live_elements"]}), flush=True)
    print("summary (element ms: min / mean over rounds; nodal ms mean)")
    for k, v in res.items():
        a = np.array(v)
        print(f"  {k:32s} {a[:, 0].min():.3f} / {a[:, 0].mean():.3f}   nodal {a[:, 1].mean():.3f}   frac(best) "
              f"{1904 * nE / (a[:, 0].min() * 1e-3) / 1e9 / 6448.4:.3f}")


if __name__ == "__main__":
    main()
