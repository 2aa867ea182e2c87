"""Error behaviour of the C ABI (argument / call-order checks), exercised on the host-compiled build of the
engine sources (tests/emu) — the CUDA library shares this code but needs a device to create an engine."""
import ctypes as C

import numpy as np
import pytest

from hakai_fem_b200.engine import HakaiError, HkParams
from hakai_fem_b200.mesh import StretchDeck
from hakai_fem_b200.model_setup import prepare, configure_engine

from .emu.emu_engine import EmuEngine, load


def test_params_size_and_dt_checked():
    lib = load()
    p = HkParams()
    lib.hke_default_params(C.byref(p))
    h = C.c_void_p()
    p.d_time = 1e-7
    p.struct_size = 12
    assert lib.hke_create(C.byref(h), C.byref(p)) == -1          # HK_ERR_ARG
    with pytest.raises(HakaiError):
        EmuEngine(d_time=0.0)
    with pytest.raises(HakaiError) as ei:
        EmuEngine(d_time=1e-7, triax_route=1)
    assert "oracle" in str(ei.value)


def test_call_order_and_ranges():
    st = prepare(StretchDeck(2, 2, 2).build_model())
    e = EmuEngine(d_time=st.d_time)
    with pytest.raises(HakaiError):
        e.finalize()                                             # no mesh yet
    with pytest.raises(HakaiError):
        e.step(1, 1)                                             # not finalised
    m = st.model
    bad = m.elementmat.copy()
    bad[0, 0] = m.nNode + 1
    with pytest.raises(HakaiError):
        e.set_mesh(m.coordmat, bad, m.element_material, m.element_instance, st.diag_M)
    dm = st.diag_M.copy()
    dm[1] *= 2
    with pytest.raises(HakaiError) as ei:
        e.set_mesh(m.coordmat, m.elementmat, m.element_material, m.element_instance, dm)
    assert "same mass" in str(ei.value)
    e.set_mesh(m.coordmat, m.elementmat, m.element_material, m.element_instance, st.diag_M)
    with pytest.raises(HakaiError):
        e.add_material(1.0, 0.3, 1.0, plastic=np.array([[1.0, 0.0]]), Hd=None)           # one-row *Plastic table
    with pytest.raises(HakaiError):
        e.add_material(1.0, 0.3, 1.0, plastic=np.array([[1.0, 0.0], [2.0, 0.0]]), Hd=np.array([1.0]))   # not increasing
    e.add_material(210000.0, 0.3, 7.8e-9)
    e.add_bc([np.array([3 * m.nNode + 1])], [0.0])               # dof out of range: reported at finalize
    with pytest.raises(HakaiError):
        e.finalize()


def test_finalize_twice_and_halo_multi_step_rejected():
    st = prepare(StretchDeck(2, 2, 2).build_model())

    def with_halo(**p):
        e = EmuEngine(**p)
        e.set_halo([np.array([1, 2, 3])])
        return e
    e = configure_engine(with_halo, st)
    with pytest.raises(HakaiError):
        e.finalize()
    with pytest.raises(HakaiError):
        e.step(1, 1)                                             # halo buffers not bound
    send, recv = np.zeros(9), np.zeros(9)
    e.halo_bind(0, send.ctypes.data, recv.ctypes.data)
    e.halo_pack()
    e.step(1, 1)
    with pytest.raises(HakaiError):
        e.step(2, 2)                                             # with halos: one step per exchange


def test_restart_and_multi_domain_hooks_are_checked():
    st = prepare(StretchDeck(2, 2, 2).build_model())
    e = EmuEngine(d_time=st.d_time)
    for call in (lambda: e.node_output(), lambda: e.mark_frame(), lambda: e.apply_deleted([1]),
                 lambda: e.set_global_maps([1], [1], [1])):
        with pytest.raises(HakaiError):
            call()                                               # not finalised
    e = configure_engine(EmuEngine, st)
    nE, nN = st.model.nElement, st.model.nNode
    with pytest.raises(HakaiError):
        e.apply_deleted([nE + 1])                                # local id out of range
    with pytest.raises(HakaiError):
        e.apply_deleted([0])
    e.apply_deleted([2])                                         # restart hook: recorded as already deleted
    assert list(e.deleted_ids()) == [2] and e.sync() == 0
    with pytest.raises(HakaiError):
        e.set_global_maps(np.full(3, nN + 1), np.ones(2), np.ones(2))       # local node id beyond the mesh
    with pytest.raises(HakaiError):
        e.set_global_maps(np.ones(3), np.full(2, nE + 1), np.ones(2))
    e.set_global_maps(np.arange(1, nN + 1), np.arange(1, nE + 1), np.ones(nE))
    with pytest.raises(HakaiError):
        e.apply_deleted([nE + 1])                                # now a GLOBAL id: still range-checked
    with pytest.raises(ValueError):
        e.node_output(out=dict(node_stress=np.zeros((nN, 6))))   # wrong layout: (6, nNode) expected
    e.mark_frame()
    e.step_enqueue(1, 1)
    assert e.sync() == 0


def test_lengths_ids_and_null_tables_are_validated():
    """the ABI must not trust lengths and ids.  A *Boundary block mixing numbered lines with ENCASTRE yields
    more dof lists than values (the reference raises BoundsError there): refused before anything is read past the end;
    contact node / element ids and instance face tables are range-checked; material tables may not be NULL."""
    import ctypes as C
    st = prepare(StretchDeck(2, 2, 2).build_model())
    e = EmuEngine(d_time=st.d_time)
    m = st.model
    e.set_mesh(m.coordmat, m.elementmat, m.element_material, m.element_instance, st.diag_M)
    with pytest.raises(HakaiError):
        e.add_bc([np.array([1, 2]), np.array([3])], [0.0])                  # 2 lists, 1 value
    with pytest.raises(HakaiError):
        e.add_ic([np.array([1, 2]), np.array([3])], [0.0])
    nN, nE = m.nNode, m.nElement
    tri = np.array([[1, 2, 3]])
    with pytest.raises(HakaiError):
        e.add_contact_pair(1, 1, [1, nN + 1], [1], tri, [1], 1.0)           # slave node beyond the mesh
    with pytest.raises(HakaiError):
        e.add_contact_pair(1, 1, [1], [1], np.array([[1, 2, 0]]), [1], 1.0)  # triangle node 0
    with pytest.raises(HakaiError):
        e.add_contact_pair(1, 1, [1], [1], tri, [nE + 1], 1.0)              # triangle element beyond the mesh
    faces = np.ones((6 * nE, 4), np.int64)
    bad = faces.copy()
    bad[3, 2] = nN + 5
    with pytest.raises(HakaiError):
        e.add_instance(0, nN, 0, nE, bad, np.repeat(np.arange(1, nE + 1), 6))
    with pytest.raises(HakaiError):
        e.add_instance(0, nN, 0, nE, faces, np.full(6 * nE, nE + 1))
    fn = e._fn("add_material")                                              # npp = 2 with NULL tables, straight at the ABI
    rc = fn(e._h, C.c_double(1.0), C.c_double(0.3), C.c_double(1.0), C.c_int64(2), None, None, C.c_int64(0), None)
    assert rc != 0 and b"NULL" in e._fn("last_error")(e._h)


def test_comm_erosion_call_order_and_capacity_are_checked():
    """hk_comm_erosion: after hk_set_global_maps, before the first step, with a sane capacity; and a deck that cannot erode
    (no contact, no failing material) simply keeps the plain path."""
    st = prepare(StretchDeck(2, 2, 2).build_model())
    e = EmuEngine(d_time=st.d_time)
    with pytest.raises(HakaiError):
        e.comm_erosion(16)                                       # not finalised
    e = configure_engine(EmuEngine, st)
    nE, nN = st.model.nElement, st.model.nNode
    with pytest.raises(HakaiError) as ei:
        e.comm_erosion(16)                                       # no global maps yet
    assert "hk_set_global_maps" in str(ei.value)
    e.set_global_maps(np.arange(1, nN + 1), np.arange(1, nE + 1), np.ones(nE))
    for bad in (0, -3, 1 << 25):
        with pytest.raises(HakaiError):
            e.comm_erosion(bad)
    e.comm_erosion(16)
    e.step_enqueue(1, 2)                                         # nothing to erode: steps run, no device lists are built
    assert e.sync() == 0
    with pytest.raises(HakaiError) as ei:
        e.comm_erosion(16)                                       # too late
    assert "first step" in str(ei.value)
    e.apply_deleted([1])                                         # the host replay still works on such an engine


def test_host_driver_rejects_unknown_contact_setup(tmp_path):
    from hakai_fem_b200.host import hakai
    path = tmp_path / "d.inp"
    StretchDeck(2, 2, 2, n_steps=4).write_inp(str(path))
    with pytest.raises(ValueError):
        hakai(str(path), str(tmp_path / "o"), engine_cls=EmuEngine, contact_setup="gpu", verbose=False)
