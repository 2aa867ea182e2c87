"""Host driver (Python twin of hakai()): frame cadence and VTK layout, driven with the CPU oracle engine."""
import os

import numpy as np

from hakai_fem_b200.host import hakai
from hakai_fem_b200.mesh import StretchDeck
from oracle.oracle_engine import OracleEngine


def test_hakai_writes_reference_style_frames(tmp_path):
    deck = StretchDeck(2, 2, 3, n_steps=200.5, strain_per_step=3e-4)
    path = tmp_path / "d.inp"
    deck.write_inp(str(path))
    out = tmp_path / "temp"
    eng, frames = hakai(str(path), str(out), engine_cls=OracleEngine, verbose=False)
    assert len(frames) == 101                                  # file000 + one per d_out = floor(200.5/100) = 2 steps
    assert os.path.basename(frames[0]) == "file000.vtk" and os.path.basename(frames[-1]) == "file100.vtk"
    txt = open(frames[-1]).read().split("\n")
    assert txt[0] == "# vtk DataFile Version 2.0" and txt[3] == "DATASET UNSTRUCTURED_GRID"
    nN = 3 * 3 * 4
    assert txt[4] == f"POINTS {nN} float"
    heads = [l for l in txt if l.startswith(("SCALARS", "VECTORS", "CELLS", "CELL_TYPES", "POINT_DATA"))]
    names = [h.split()[1] for h in heads if h.startswith(("SCALARS", "VECTORS"))]
    assert names == ["DISPLACEMENT", "Vx", "Vy", "Vz", "E11", "E22", "E33", "E12", "E23", "E13", "EQ_PSTRAIN",
                     "S11", "S22", "S33", "S12", "S23", "S13", "MISES_STRESS", "TRIAX_STRESS"]
    assert f"CELLS {12} {12 * 9}" in txt
    d = eng.download()
    i = txt.index("VECTORS DISPLACEMENT float")
    row = np.array(txt[i + nN].split(), float)                 # last node
    assert np.allclose(row, d["disp"][-3:], rtol=1e-6, atol=1e-12)


def _read_binary_vtk(path):
    """Minimal legacy-VTK BINARY reader for the sections write_vtk emits."""
    raw = open(path, "rb").read()
    pos = 0
    out = {}

    def line():
        nonlocal pos
        end = raw.index(b"\n", pos)
        s = raw[pos:end].decode()
        pos = end + 1
        return s

    def take(n, dt):
        nonlocal pos
        a = np.frombuffer(raw, dt, n, pos)
        pos += a.nbytes + 1                                    # + newline
        return a
    assert [line() for _ in range(4)] == ["# vtk DataFile Version 2.0", "Test", "BINARY", "DATASET UNSTRUCTURED_GRID"]
    n = int(line().split()[1])
    out["POINTS"] = take(3 * n, ">f4").reshape(n, 3)
    nc, tot = map(int, line().split()[1:])
    out["CELLS"] = take(tot, ">i4").reshape(nc, 9)
    assert int(line().split()[1]) == nc
    out["CELL_TYPES"] = take(nc, ">i4")
    assert int(line().split()[1]) == n
    assert line() == "VECTORS DISPLACEMENT float"
    out["DISPLACEMENT"] = take(3 * n, ">f4").reshape(n, 3)
    names = []
    while pos < len(raw):
        h = line().split()
        assert h[0] == "SCALARS" and line() == "LOOKUP_TABLE default"
        out[h[1]] = take(n, ">f4")
        names.append(h[1])
    return out, names


def test_binary_frames_and_device_node_output_match_the_reference_route(tmp_path):
    """Same run three ways: reference route (host averaging, ASCII), device averaging + ASCII, device + BINARY."""
    deck = StretchDeck(3, 2, 4, n_steps=100, strain_per_step=4e-4, jitter=0.05)
    path = tmp_path / "d.inp"
    deck.write_inp(str(path))
    runs = {}
    for tag, kw in (("ref", dict(node_output="host")), ("dev", {}), ("bin", dict(vtk_format="binary"))):
        _, frames = hakai(str(path), str(tmp_path / tag), engine_cls=OracleEngine, verbose=False, **kw)
        runs[tag] = frames
    for a, b in zip(runs["ref"], runs["dev"]):                 # %1.6e text: identical unless a digit sits on a tie
        ta, tb = open(a).read().split("\n"), open(b).read().split("\n")
        assert len(ta) == len(tb)
        diff = [i for i, (x, y) in enumerate(zip(ta, tb)) if x != y]
        for i in diff:
            assert np.allclose(np.array(ta[i].split(), float), np.array(tb[i].split(), float), rtol=2e-6, atol=1e-15)
        assert len(diff) <= len(ta) // 100
    last_txt = open(runs["dev"][-1]).read().split("\n")
    got, names = _read_binary_vtk(runs["bin"][-1])
    assert names == ["Vx", "Vy", "Vz", "E11", "E22", "E33", "E12", "E23", "E13", "EQ_PSTRAIN",
                     "S11", "S22", "S33", "S12", "S23", "S13", "MISES_STRESS", "TRIAX_STRESS"]
    nN = 4 * 3 * 5
    assert (got["CELL_TYPES"] == 12).all() and (got["CELLS"][:, 0] == 8).all()
    for name in ("S33", "EQ_PSTRAIN", "MISES_STRESS"):
        i = last_txt.index(f"SCALARS {name} float 1") + 2
        want = np.array(last_txt[i:i + nN], float)
        assert np.abs(want).max() > 0
        assert np.allclose(got[name], want, rtol=2e-6, atol=1e-12), name
    i = last_txt.index("VECTORS DISPLACEMENT float") + 1
    want = np.array([l.split() for l in last_txt[i:i + nN]], float)
    assert np.allclose(got["DISPLACEMENT"], want, rtol=2e-6, atol=1e-12)
    assert os.path.getsize(runs["bin"][-1]) < 0.5 * os.path.getsize(runs["dev"][-1])


def test_resume_from_checkpoint_writes_identical_frames(tmp_path):
    deck = StretchDeck(2, 2, 3, n_steps=100, strain_per_step=4e-4, jitter=0.05)
    path = tmp_path / "d.inp"
    deck.write_inp(str(path))
    ck = str(tmp_path / "ck.npz")
    _, full = hakai(str(path), str(tmp_path / "a"), engine_cls=OracleEngine, verbose=False, checkpoint=ck,
                    checkpoint_frames=40)                      # written at frame 40 and 80; the last one survives
    _, tail = hakai(str(path), str(tmp_path / "b"), engine_cls=OracleEngine, verbose=False, resume=ck)
    assert [os.path.basename(f) for f in tail] == ["file%03d.vtk" % i for i in range(81, 101)]
    for f in tail:
        assert open(f).read() == open(os.path.join(str(tmp_path / "a"), os.path.basename(f))).read()
