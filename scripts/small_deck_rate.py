"""Steps per second of the reference's small example decks on the CUDA engine (launch-bound regime):
python scripts/small_deck_rate.py [deck ...]   — prints microseconds per step (wall clock around one hk_step call)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from hakai_fem_b200.engine import Engine
    from hakai_fem_b200.model_setup import configure_engine, prepare
    from tests import util
    names = sys.argv[1:] or ["t5", "bullet_impact", "metal_cutting", "car_crash_n2k"]
    for name in names:
        st = prepare(util.t5_model()) if name == "t5" else util.deck_setup(name)
        g = configure_engine(Engine, st)
        g.step(1, 200)                         # warm-up (module load, first-step set-up)
        n = 3000
        t0 = time.perf_counter()
        g.step_enqueue(201, n)
        t_enq = time.perf_counter() - t0
        g.sync()
        dt = time.perf_counter() - t0
        c = g.counters()
        print(f"{name:16s} nE {st.model.nElement:7d} contact {int(st.model.contact_flag)}  {dt / n * 1e6:7.1f} us/step "
              f"(host enqueue {t_enq / n * 1e6:6.1f} us/step, {int(c[3])} launches so far)", flush=True)
        g.close()


if __name__ == "__main__":
    main()
