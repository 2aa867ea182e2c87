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
