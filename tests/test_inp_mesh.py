"""Deck reader (Python mirror of readInpFile_j.jl) and synthetic deck generator."""
import os

import numpy as np
import pytest

from hakai_fem_b200.inp import read_inp_file, parse_inp_lines
from hakai_fem_b200.mesh import StretchDeck, steel
from hakai_fem_b200.model_setup import prepare

from . import util


@pytest.mark.skipif(not os.path.exists(util.REF_T5), reason="reference tree not mounted")
def test_tensile5e_parse_matches_committed_fixture():
    m = read_inp_file(util.REF_T5)
    f = util.t5_model()
    assert (m.nNode, m.nElement) == (24, 5)
    assert np.array_equal(m.coordmat, f.coordmat) and np.array_equal(m.elementmat, f.elementmat)
    assert np.array_equal(m.element_material, f.element_material)
    assert (m.d_time, m.end_time, m.mass_scaling, m.contact_flag) == (f.d_time, f.end_time, f.mass_scaling, 0)
    for a, b in zip(m.BC, f.BC):
        assert a.amp_name == b.amp_name and a.value == b.value
        assert all(np.array_equal(x, y) for x, y in zip(a.dof, b.dof))


def test_tensile5e_closed_form_facts():
    """SURVEY §8c anchors: 5 bricks 10x10x5 mm."""
    st = prepare(util.t5_model())
    m = st.model
    assert np.allclose(st.elementVolume, 500.0, rtol=1e-14)
    assert np.isclose(st.diag_M.sum() / 3, 1.95e-5, rtol=1e-13)
    assert np.isclose(st.diag_M.min(), 4.875e-7, rtol=1e-13)
    assert (st.elementMinSize, st.elementMaxSize) == (5.0, 10.0)
    assert st.time_num == 20000.0
    mat = m.MATERIAL[m.element_material[0] - 1]
    assert (mat.young, mat.poisson, mat.plastic[0, 0], mat.plastic.shape[0]) == (210000.0, 0.3, 755.0, 8)
    assert np.array_equal(mat.ductile, [[1.0, 0.0, 30.0], [0.3, 0.3, 30.0]])
    fixed = sorted(set(((m.BC[0].dof[0] - 1) // 3 + 1).tolist()))
    assert fixed == [1, 7, 13, 19]
    driven = sorted(set(((m.BC[1].dof[1] - 1) // 3 + 1).tolist()))
    assert driven == [6, 12, 18, 24] and m.BC[1].value[1] == 10.0
    # only the last line of an *Amplitude block survives (R2:649-665); T5 has one line with two points
    assert np.array_equal(m.BC[1].amplitude.time, [0.0, 0.01]) and np.array_equal(m.BC[1].amplitude.value, [0.0, 1.0])


QUIRK_DECK = """*Heading
*Part, name=P
*Node
1, 0., 0., 0.
2, 1., 0., 0.
3, 1., 1., 0.
4, 0., 1., 0.
5, 0., 0., 1.
6, 1., 0., 1.
7, 1., 1., 1.
8, 0., 1., 1.
*Element, type=C3D8R
1, 1, 2, 3, 4, 5, 6, 7, 8
*Nset, nset=all, generate
1, 8, 1
*Nset, nset=ignored
1, 2
*Solid Section, elset=all, material=M
,
*End Part
*Assembly, name=A
*Instance, name=I1, part=P
 1., 2., 3.
*End Instance
*Nset, nset=top, instance=I1
 5, 6, 7,
 8
*Nset, nset=bot, instance=I1, generate
 1, 4, 1
*End Assembly
*Amplitude, name=Amp
 0., 0., 1., 5.
 2., 7., 3., 9.
*Material, name=M
*Density
 2.0,
*Elastic
 100., 0.25
**
*Boundary
bot, ENCASTRE
**
*Boundary, amplitude=Amp
top, 1, 1
top, 3, 3, 0.5
top, 5, 5
**
*Initial Conditions, type=VELOCITY
I1.all, 2, -3.
**
*Step
*Dynamic, Explicit
1e-3, 0.1
*Fixed Mass Scaling, factor=4.
*End Step
"""


def test_reader_quirks():
    m = parse_inp_lines(QUIRK_DECK.split("\n"))
    assert m.nNode == 8 and m.nElement == 1
    assert np.allclose(m.coordmat[:, 0], [1.0, 2.0, 3.0])                      # instance translation
    assert [n.name for n in m.PART[0].NSET] == ["all"]                         # part-level Nset only with generate
    assert np.array_equal(m.NSET[0].nodes, [5, 6, 7, 8]) and np.array_equal(m.NSET[1].nodes, [1, 2, 3, 4])
    assert np.array_equal(m.AMPLITUDE[0].time, [2.0, 3.0]) and np.array_equal(m.AMPLITUDE[0].value, [7.0, 9.0])
    enc = m.BC[0]
    assert enc.value == [0.0] and np.array_equal(enc.dof[0], [1, 4, 7, 10, 2, 5, 8, 11, 3, 6, 9, 12])
    drv = m.BC[1]
    assert drv.amp_name == "Amp" and drv.value == [0.0, 0.5] and len(drv.dof) == 2      # direction 5 dropped
    assert np.array_equal(drv.dof[1], np.array([5, 6, 7, 8]) * 3)
    assert np.array_equal(m.IC[0].dof[0], np.arange(1, 9) * 3 - 1) and m.IC[0].value == [-3.0]
    assert (m.d_time, m.end_time, m.mass_scaling) == (1e-3, 0.1, 4.0)
    st = prepare(m)
    assert np.isclose(st.d_time, 2e-3) and np.isclose(st.diag_M[0], 2.0 * 1.0 / 8 * 4.0)


def test_written_deck_equals_direct_arrays(tmp_path):
    deck = StretchDeck(3, 2, 4, jitter=0.1, n_steps=49.5,
                       material=steel("steel_Ductile", ductile=[[0.03, 0.0, 30.0], [0.02, 0.3, 30.0]]))
    path = tmp_path / "deck.inp"
    deck.write_inp(str(path))
    a, b = read_inp_file(str(path)), deck.build_model()
    assert np.array_equal(a.coordmat, b.coordmat) and np.array_equal(a.elementmat, b.elementmat)
    assert (a.d_time, a.end_time) == (b.d_time, b.end_time)
    assert np.array_equal(a.MATERIAL[0].plastic, b.MATERIAL[0].plastic)
    assert np.array_equal(a.MATERIAL[0].ductile, b.MATERIAL[0].ductile)
    assert len(a.BC) == 1 and a.BC[0].value == b.BC[0].value
    assert all(np.array_equal(x, y) for x, y in zip(a.BC[0].dof, b.BC[0].dof))
    assert a.IC[0].value == b.IC[0].value and all(np.array_equal(x, y) for x, y in zip(a.IC[0].dof, b.IC[0].dof))
    assert np.array_equal(a.BC[0].amplitude.time, b.BC[0].amplitude.time)


def test_written_impact_deck_equals_direct_arrays(tmp_path):
    """ImpactDeck.write_inp (two parts, translated instance, assembly-level set, *Contact ... ALL EXTERIOR) read back
    by the readInpFile mirror gives the arrays of build_model(), down to the contact surfaces and the lumped mass."""
    from hakai_fem_b200.mesh import ImpactDeck
    from hakai_fem_b200.model_setup import prepare
    deck = ImpactDeck(plate=(6, 5, 2), proj=(2, 3, 2), v0=-300.0, plate_ductile=[[0.02, 0.0, 30.0], [0.015, 0.4, 30.0]])
    path = tmp_path / "impact.inp"
    deck.write_inp(str(path))
    a, b = read_inp_file(str(path)), deck.build_model()
    assert (a.nNode, a.nElement, a.contact_flag, a.d_time, a.end_time) == (b.nNode, b.nElement, 1, b.d_time, b.end_time)
    assert np.array_equal(a.coordmat, b.coordmat) and np.array_equal(a.elementmat, b.elementmat)
    assert np.array_equal(a.element_material, b.element_material)
    assert np.array_equal(a.element_instance, b.element_instance)
    assert [m.name for m in a.MATERIAL] == ["alum", "lead"]
    assert np.array_equal(a.MATERIAL[0].ductile, b.MATERIAL[0].ductile)
    assert np.array_equal(np.sort(np.concatenate(a.BC[0].dof)), np.sort(np.concatenate(b.BC[0].dof)))
    assert a.IC[0].value == b.IC[0].value and np.array_equal(a.IC[0].dof[0], b.IC[0].dof[0])
    sa, sb = prepare(a), prepare(b)
    assert np.array_equal(sa.diag_M, sb.diag_M) and len(sa.CT) == len(sb.CT) == 2
    for x, y in zip(sa.CT, sb.CT):
        assert (x.i_instance, x.j_instance, x.young) == (y.i_instance, y.j_instance, y.young)
        assert np.array_equal(x.c_triangles, y.c_triangles) and np.array_equal(x.c_nodes_i, y.c_nodes_i)
        assert np.array_equal(x.c_nodes_j, y.c_nodes_j) and np.array_equal(x.c_triangles_eleid, y.c_triangles_eleid)
