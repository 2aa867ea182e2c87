"""Writes the benchmark decks of SURVEY §8d as real Abaqus `.inp` files, so the SAME input can be fed to the reference
(`julia HAKAI_j.jl deck.inp`) wherever Julia exists and to `python -m hakai_fem_b200.host deck.inp`.

    python scripts/make_decks.py outdir [B1] [F16] [I8] [--scale 0.25]

B1: 50x50x400 elastoplastic bar (99 steps, frame-free as SURVEY prescribes); F16: 252^3 ductile block with jittered
nodes; I8: 400x400x48 plate + 68^3 projectile, ALL EXTERIOR contact.  --scale shrinks every edge count (smoke sizes).
T5 is the reference's own HAKAI-v0.0.0/input/Tensile5e.inp and is not regenerated."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hakai_fem_b200.mesh import B1, F16, ImpactDeck            # noqa: E402


def main(argv):
    if not argv:
        raise SystemExit(__doc__)
    outdir, names, scale = argv[0], [], 1.0
    it = iter(argv[1:])
    for a in it:
        if a == "--scale":
            scale = float(next(it))
        else:
            names.append(a)
    names = names or ["B1", "F16", "I8"]
    os.makedirs(outdir, exist_ok=True)
    s = lambda n: max(1, int(round(n * scale)))
    for name in names:
        path = os.path.join(outdir, f"{name}.inp")
        if name == "B1":
            deck = B1(scale=scale, n_steps=99.5)
        elif name == "F16":
            deck = F16(n=s(252), n_steps=99.5)
        elif name == "I8":
            deck = ImpactDeck(plate=(s(400), s(400), s(48)), proj=(s(68), s(68), s(68)))
        else:
            raise SystemExit(f"unknown deck {name}")
        deck.write_inp(path)
        print(name, path, "%.1f MB" % (os.path.getsize(path) / 1e6))


if __name__ == "__main__":
    main(sys.argv[1:])
