"""Parity cases shared by the CPU (host-emulated kernel bodies) and GPU (CUDA engine) test modules.

Every case drives the engine under test and the CPU oracle through the same C ABI on the same inputs.
Tolerances (FP64, relative to the field's max magnitude):
  TOL_STEP  = 1e-13  one step from identical state: only FMA contraction / summation-form differences
  TOL_RUN   = 1e-9   after N <= 20 000 steps on Tensile5e
Integer results (element flags, deleted ids and their order, contact surface lists) must be identical.
"""
import numpy as np

from hakai_fem_b200.model_setup import prepare, configure_engine
from hakai_fem_b200.mesh import ImpactDeck
from oracle.oracle_engine import OracleEngine

from . import util

TOL_STEP = 1e-13
TOL_RUN = 1e-9
STATE_KEYS = ("disp", "velo", "integ_stress", "integ_strain", "integ_eq_plastic_strain", "integ_triax_stress",
              "element_flag", "disp_pre", "Q", "position", "integ_yield_stress")


def case_roundtrip(engine_cls):
    """hk_upload_state -> hk_download is the identity (layout transposes are exact)."""
    st = prepare(util.distorted_block().build_model())
    eng = configure_engine(engine_cls, st)
    rs = util.random_state(st, seed=3)
    eng.upload_state(**rs)
    d = util.full_state(eng)
    for k in ("disp", "disp_pre", "velo", "Q", "integ_stress", "integ_strain", "integ_eq_plastic_strain",
              "integ_yield_stress"):
        assert np.array_equal(np.asarray(d[k]), np.asarray(rs[k])), k
    assert np.array_equal(d["position"], st.model.coordmat + rs["disp"].reshape(-1, 3).T)


def case_single_step_random_state(engine_cls, ductile=False):
    """One step from an identical random elastic/plastic state on a distorted mesh."""
    st = prepare(util.distorted_block(ductile=ductile).build_model())
    o, g = util.make_pair(st, engine_cls, OracleEngine)
    rs = util.random_state(st, seed=11)
    o.upload_state(**rs)
    g.upload_state(**rs)
    o.step(5, 1)
    g.step(5, 1)
    a, b = util.full_state(o), util.full_state(g)
    util.assert_states_close(a, b, TOL_STEP, STATE_KEYS, "single step")
    # the nodal kernel mirrors the reference's operation order: bit-exact on identical inputs
    assert np.array_equal(a["disp"], b["disp"])
    assert np.array_equal(a["disp_pre"], b["disp_pre"])


def case_t5(engine_cls, n_total=20000):
    """Tensile5e.inp: free-running trajectories, snapshots, deletion of element 3 at step 15153."""
    st = prepare(util.t5_model())
    gold = util.load_json("tensile5e_oracle.json")
    o, g = util.make_pair(st, engine_cls, OracleEngine)
    t = 0
    for target in (1, 2, 10, 316, 1000, 5000, 15152, 15153, 20000):
        if target > n_total:
            break
        n = target - t
        nd_o = o.step(t + 1, n)
        nd_g = g.step(t + 1, n)
        t = target
        assert nd_o == nd_g, f"deletions differ in steps up to {t}"
        a, b = util.full_state(o), util.full_state(g)
        util.assert_states_close(a, b, TOL_RUN, STATE_KEYS, f"T5 step {t}")
        snap = gold["snapshots"][str(t)]
        assert util.rel_err(b["disp"], snap["disp"]) <= TOL_RUN
        assert util.rel_err(b["integ_eq_plastic_strain"], snap["eps"]) <= TOL_RUN
        assert np.array_equal(b["element_flag"], np.array(snap["element_flag"]))
    if n_total >= 20000:
        assert g.deleted_ids().tolist() == gold["deleted"] == [3]


def case_fracture_block(engine_cls, n_steps=260):
    """Jittered ductile block under uniform stretch: elements delete at different steps; the set, the
    order and the step of every deletion must match the oracle."""
    deck = util.distorted_block(nx=5, ny=4, nz=6, jitter=0.05, ductile=True, strain_per_step=4e-4)
    st = prepare(deck.build_model())
    o, g = util.make_pair(st, engine_cls, OracleEngine)
    t = 0
    seen = 0
    while t < n_steps:
        n = 7                       # batches: deletions inside a batch must still come out in order
        do, dg = o.step(t + 1, n), g.step(t + 1, n)
        t += n
        assert do == dg, f"deleted count differs at step {t}"
        seen += do
        assert np.array_equal(o.deleted_ids(), g.deleted_ids())
    assert seen > 3, "deck did not delete anything: test is vacuous"
    a, b = util.full_state(o), util.full_state(g)
    util.assert_states_close(a, b, 1e-8, STATE_KEYS, "fracture block")
    so, sg = o.deleted_steps(), g.deleted_steps()          # the step of every deletion, aligned with deleted_ids
    assert np.array_equal(so, sg) and len(sg) == seen and np.all(np.diff(sg) >= 0) and sg.min() >= 1 and sg.max() <= t


def case_state_summary(engine_cls):
    """hk_state_summary (device-side reduction) against the oracle's plain loops and against the downloaded arrays,
    before yield, in the plastic regime and after deletions."""
    deck = util.distorted_block(nx=5, ny=4, nz=6, jitter=0.05, ductile=True, strain_per_step=4e-4)
    st = prepare(deck.build_model())
    o, g = util.make_pair(st, engine_cls, OracleEngine)
    s0 = g.state_summary()
    assert s0 == dict(live_elements=st.model.nElement, eps_min=0.0, eps_max=0.0, yielded_points=0)
    t = 0
    for n in (30, 31):                   # step 61 deletes 117 of the 120 elements, step 62 the rest
        o.step(t + 1, n)
        g.step(t + 1, n)
        t += n
        a, b = o.state_summary(), g.state_summary()
        d = g.download(fields=("integ_eq_plastic_strain", "element_flag"))
        live = np.repeat(d["element_flag"] == 1, 8)
        eps = d["integ_eq_plastic_strain"][live]
        assert b["live_elements"] == int((d["element_flag"] == 1).sum()) == a["live_elements"]
        assert b["eps_min"] == eps.min() and b["eps_max"] == eps.max() and b["yielded_points"] == int((eps > 0).sum())
        assert a["yielded_points"] == b["yielded_points"]
        assert abs(a["eps_max"] - b["eps_max"]) <= 1e-9 * max(a["eps_max"], 1e-30)
    assert b["live_elements"] < st.model.nElement, "deck did not delete anything: test is vacuous"
    assert b["eps_min"] > 0


def case_node_output(engine_cls):
    """hk_node_output = cal_node_stress_strain (J2:3408-3486) on the device: bit-identical to the oracle's loop on an
    identical ip state (same summation orders), equal to the NumPy host twin up to summation order, and consistent
    with it after a real run with deletions."""
    from hakai_fem_b200.host import cal_node_stress_strain
    deck = util.distorted_block(nx=5, ny=4, nz=6, jitter=0.05, ductile=True, strain_per_step=4e-4)
    st = prepare(deck.build_model())
    o, g = util.make_pair(st, engine_cls, OracleEngine)
    rs = util.random_state(st, seed=5)
    o.upload_state(**rs)
    g.upload_state(**rs)
    a, b = o.node_output(), g.node_output()
    for k in ("node_stress", "node_strain", "node_eq_plastic_strain", "node_mises_stress", "inc_num"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(b["inc_num"], np.bincount(st.model.elementmat.reshape(-1) - 1, minlength=st.model.nNode))
    raw = g.node_output(raw=True)
    assert np.array_equal(raw["node_stress"] / raw["inc_num"][:, None], b["node_stress"])
    assert "node_mises_stress" not in raw
    nN = st.model.nNode
    bufs = dict(node_stress=np.empty((6, nN)), node_mises_stress=np.empty(nN))      # caller-owned (e.g. pinned) buffers
    c = g.node_output(out=bufs)
    assert np.shares_memory(c["node_stress"], bufs["node_stress"]) and c["node_mises_stress"] is bufs["node_mises_stress"]
    assert np.array_equal(c["node_stress"], b["node_stress"]) and np.array_equal(c["node_mises_stress"], b["node_mises_stress"])
    o.step(1, 200)                                    # with deletions (zeroed stress of dead elements is averaged in)
    nd = g.step(1, 200)
    assert nd > 0
    a, b = o.node_output(), g.node_output()
    h = cal_node_stress_strain(st.model.nNode, st.model.elementmat, 8, g.download())
    for k in ("node_stress", "node_strain", "node_eq_plastic_strain", "node_mises_stress", "node_triax_stress"):
        assert util.rel_err(h[k], b[k]) <= 1e-13, f"host twin {k}: {util.rel_err(h[k], b[k])}"
        assert util.rel_err(a[k], b[k]) <= 1e-8, f"oracle {k}: {util.rel_err(a[k], b[k])}"


# SI units, E = 7e10 Pa: anything below these magnitudes is rounding noise of a body in rigid motion
CONTACT_FLOORS = dict(integ_stress=1e4, integ_strain=1e-7, Q=1e-3, integ_yield_stress=1.0, external_force=1e-3)


def small_impact(mu=0.25, plate=(8, 8, 2), proj=(3, 3, 3)):
    deck = ImpactDeck(plate=plate, proj=proj)
    st = prepare(deck.build_model())
    return st, dict(contact_myu=mu)


def case_contact(engine_cls, mu=0.25, n_steps=60):
    """Two-instance impact: contact forces, hit counts and trajectories vs the oracle."""
    st, prm = small_impact(mu)
    o, g = util.make_pair(st, engine_cls, OracleEngine, **prm)
    t = 0
    max_f = 0.0
    for _ in range(n_steps // 4):
        o.step(t + 1, 4)
        g.step(t + 1, 4)
        t += 4
        a, b = util.full_state(o), util.full_state(g)
        max_f = max(max_f, float(np.abs(a["external_force"]).max()))
        # contact sums: exact fixed-point accumulation vs Float128 -> same doubles up to rare 1-ulp ties
        # before the bodies touch, stresses are pure rounding noise (rigid motion): floors give the scale
        keys = tuple(k for k in STATE_KEYS if k != "integ_triax_stress") + ("external_force",)
        util.assert_states_close(a, b, 1e-9, keys, f"contact step {t}", floors=CONTACT_FLOORS)
        tot = a["external_force"].reshape(-1, 3).sum(axis=0)
        assert np.all(np.abs(tot) <= 1e-9 * max(max_f, 1e-300)), "net contact force must vanish (J2:2653-2667)"
    assert max_f > 0, "no contact happened: test is vacuous"
    co, cg = o.counters(), g.counters()
    assert co[1] == cg[1] and co[1] > 0, f"hit counts differ: {co[1]} vs {cg[1]}"


def case_contact_clamp(engine_cls, n_steps=60):
    """v0.0.1's penetration-rate clamp (hk_params.contact_dmax_clamp; HAKAI-v0.0.1 HAKAI_j.jl:2756, 2898, 618): a slave
    node's penetration may grow by at most d_max = max |d_disp| of the previous step.  Engine vs oracle with the clamp
    on; and the clamp must actually bite (forces differ from the v0.0.2 run of the same deck)."""
    st, prm = small_impact(0.25)
    o, g = util.make_pair(st, engine_cls, OracleEngine, contact_dmax_clamp=1, **prm)
    free = configure_engine(OracleEngine, st, **prm)
    keys = tuple(k for k in STATE_KEYS if k != "integ_triax_stress") + ("external_force",)
    t, differs = 0, False
    for _ in range(n_steps // 4):
        o.step(t + 1, 4)
        g.step(t + 1, 4)
        free.step(t + 1, 4)
        t += 4
        a, b = util.full_state(o), util.full_state(g)
        util.assert_states_close(a, b, 1e-9, keys, f"clamp step {t}", floors=CONTACT_FLOORS)
        f0 = free.download_ex(fields=("external_force",))["external_force"]
        differs = differs or util.rel_err(f0, a["external_force"]) > 1e-3
    assert o.counters()[1] == g.counters()[1] > 0
    assert differs, "the clamp never changed a force: test is vacuous"
    # one call for all steps (the d_node / d_max ping-pong is enqueued, not driven by host reads)
    g2 = configure_engine(engine_cls, st, contact_dmax_clamp=1, **prm)
    g2.step(1, t)
    assert np.array_equal(g2.download()["disp"], b["disp"])


def case_contact_single_step(engine_cls):
    """Contact force of ONE step from an identical penetrated state is bit-for-bit the oracle's
    (hk_exact.cu mirrors the reference's operation order; sums are exact)."""
    st, prm = small_impact(0.25)
    o, g = util.make_pair(st, engine_cls, OracleEngine, **prm)
    o.step(1, 24)
    s = util.full_state(o)
    up = dict(disp=s["disp"], disp_pre=s["disp_pre"], velo=s["velo"], Q=s["Q"], integ_stress=s["integ_stress"],
              integ_strain=s["integ_strain"], integ_eq_plastic_strain=s["integ_eq_plastic_strain"],
              integ_yield_stress=s["integ_yield_stress"])
    g.upload_state(**up)
    o.step(25, 1)
    g.step(25, 1)
    a, b = util.full_state(o), util.full_state(g)
    assert np.abs(a["external_force"]).max() > 0
    assert util.rel_err(a["external_force"], b["external_force"]) <= 1e-15
    assert np.array_equal(a["disp"], b["disp"])


def case_contact_erosion(engine_cls, n_steps=400):
    """Impact with a brittle plate: deletions expose interior faces, which must join the contact surface
    exactly as add_surface_triangle does (J2:767-804)."""
    deck = ImpactDeck(plate=(8, 8, 2), proj=(3, 3, 3), v0=-900.0)
    model = deck.build_model()
    model.MATERIAL[0].ductile = np.array([[0.02, 0.0, 30.0], [0.015, 0.4, 30.0]])
    st = prepare(model)
    o, g = util.make_pair(st, engine_cls, OracleEngine)
    t = 0
    while t < n_steps:
        do, dg = o.step(t + 1, 10), g.step(t + 1, 10)
        t += 10
        assert do == dg, f"deleted count differs at step {t}: {do} vs {dg}"
    ids = o.deleted_ids()
    assert len(ids) > 0, "nothing eroded: test is vacuous"
    assert np.array_equal(ids, g.deleted_ids())
    for c in range(2):
        po, pg = o.contact_pair(c), g.contact_pair(c)
        for k in ("c_nodes_i", "c_nodes_j", "c_triangles", "c_triangles_eleid"):
            assert np.array_equal(po[k], pg[k]), f"pair {c} {k}"
    a, b = util.full_state(o), util.full_state(g)
    util.assert_states_close(a, b, 1e-7, ("disp", "integ_eq_plastic_strain", "element_flag"), "erosion")
    # the exposed-face update runs on the device (hk_erode_kernel), so ALL steps can be enqueued in one call without the
    # host looking at the deletions in between: bit-identical to the run above that synchronised every 10 steps
    g2 = configure_engine(engine_cls, st)
    g2.step_enqueue(1, n_steps)
    assert g2.sync() == len(ids)
    assert np.array_equal(ids, g2.deleted_ids())
    for c in range(2):
        pg, p2 = g.contact_pair(c), g2.contact_pair(c)
        for k in pg:
            assert np.array_equal(pg[k], p2[k]), f"one-call run: pair {c} {k}"
    b2 = util.full_state(g2)
    for k in ("disp", "integ_eq_plastic_strain", "integ_stress", "element_flag", "external_force"):
        assert np.array_equal(b[k], b2[k]), f"one-call run: {k}"


def case_contact_pair_surfaces(engine_cls, tmp_path):
    """`*Contact Pair` between two `*Surface` element sets (readInpFile_j.jl:1062-1103; get_surface_triangle's elset
    filter, HAKAI_j.jl:2094-2119) through the deck reader: the contact surfaces are the exterior faces of the plate's
    top element layer and of the projectile's bottom layer only — fewer than ALL EXTERIOR — and the engine matches the
    oracle on them (forces bit-exact per step, equal hit counts)."""
    from hakai_fem_b200.inp import read_inp_file
    kw = dict(plate=(16, 16, 3), proj=(5, 5, 5))
    path = str(tmp_path / "pair.inp")
    ImpactDeck(contact_pair=True, **kw).write_inp(path)
    model = read_inp_file(path)
    assert len(model.CP) == 1 and model.contact_flag == 1 and len(model.SURFACE) == 2
    st = prepare(model)
    st_all = prepare(ImpactDeck(**kw).build_model())
    assert st.all_exterior_flag == 0 and st_all.all_exterior_flag == 1
    assert [(c.i_instance, c.j_instance) for c in st.CT] == [(2, 1), (1, 2)]
    n_tri = sorted(len(c.c_triangles_eleid) for c in st.CT)
    assert n_tri == [2 * (5 * 5 + 4 * 5), 2 * (16 * 16 + 4 * 16)]          # top/bottom face + the layer's side faces
    assert sum(n_tri) < sum(len(c.c_triangles_eleid) for c in st_all.CT)
    o, g = util.make_pair(st, engine_cls, OracleEngine, contact_myu=0.25)
    keys = tuple(k for k in STATE_KEYS if k != "integ_triax_stress") + ("external_force",)
    for t0 in (1, 21, 41):
        o.step(t0, 20)
        g.step(t0, 20)
        a, b = util.full_state(o), util.full_state(g)
        util.assert_states_close(a, b, 1e-9, keys, f"contact pair step {t0 + 19}", floors=CONTACT_FLOORS)
    assert o.counters()[1] == g.counters()[1] > 0
    for c in range(2):
        po, pg = o.contact_pair(c), g.contact_pair(c)
        for k in po:
            assert np.array_equal(po[k], pg[k]), (c, k)


def case_build_contact(engine_cls, name, n_steps):
    """hk_build_contact (contact set-up on the device: faces, orientation, exterior faces by radix sort, pair lists; A12,
    SURVEY 8f.1) against the host mirror of get_element_face / get_surface_triangle: identical node lists, triangles
    and element ids for every ordered pair of the reference's deck, and — the exposed-face table coming out of the same
    sort — identical erosion: a run on the device-built tables is bit-identical to the run on the host-built ones."""
    import copy
    st_h = util.deck_setup(name)
    st_d = copy.copy(st_h)
    st_d.contact_on_device = True
    gh, gd = configure_engine(engine_cls, st_h), configure_engine(engine_cls, st_d)
    assert len(st_h.CT) > 0
    for c in range(len(st_h.CT)):
        a, b = gh.contact_pair(c), gd.contact_pair(c)
        for k in a:
            assert np.array_equal(a[k], b[k]), (name, c, k)
    nh, nd = gh.step(1, n_steps), gd.step(1, n_steps)
    assert nh == nd and np.array_equal(gh.deleted_ids(), gd.deleted_ids())
    a, b = util.full_state(gh), util.full_state(gd)
    for k in ("disp", "integ_stress", "integ_eq_plastic_strain", "element_flag", "external_force"):
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), (name, k)
    for c in range(len(st_h.CT)):
        pa, pb = gh.contact_pair(c), gd.contact_pair(c)
        for k in pa:
            assert np.array_equal(pa[k], pb[k]), (name, "after erosion", c, k)
    return nh


def case_bc_edge_cases(engine_cls):
    """Boundary-condition corners of J2:585-617: a multi-segment amplitude table, times outside every segment (the
    search falls back to segment 1 and extrapolates it), a BC without amplitude, later BCs overriding earlier ones on
    the same dof, an empty dof list — and hk_step with n_steps = 0."""
    from hakai_fem_b200.inp import Amplitude, BC
    deck = util.distorted_block(nx=3, ny=3, nz=4, jitter=0.1)
    model = deck.build_model()
    per = 4 * 4
    top = np.arange(4 * per + 1, 5 * per + 1, dtype=np.int64)
    bottom = np.arange(1, per + 1, dtype=np.int64)
    dt = model.d_time
    amp = Amplitude(name="kink", time=np.array([0.0, 40 * dt, 80 * dt]), value=np.array([0.0, 1.0, 0.25]))
    b1 = BC(Nset_name="top", amp_name="kink", amplitude=amp)
    b1.dof, b1.value = [top * 3, top * 3 - 2], [0.02, 0.005]
    b2 = BC(Nset_name="bottom")                                  # no amplitude: amp = 1
    b2.dof, b2.value = [bottom * 3, bottom * 3 - 1, np.zeros(0, np.int64)], [0.0, 0.0, 7.0]
    b3 = BC(Nset_name="override", amp_name="kink", amplitude=amp)
    b3.dof, b3.value = [top[:4] * 3 - 2], [-0.01]                # same dofs as part of b1's second list: the later BC wins
    model.BC = [b1, b2, b3]
    model.IC = []
    st = prepare(model)
    o, g = util.make_pair(st, engine_cls, OracleEngine)
    assert o.step(1, 0) == 0 and g.step(1, 0) == 0               # nothing happens, nothing breaks
    a, b = util.full_state(o), util.full_state(g)
    assert np.array_equal(a["disp"], b["disp"]) and not a["disp"].any()
    t = 0
    for n in (39, 2, 38, 2, 40):                                 # across the kink (40), the table end (80) and beyond
        o.step(t + 1, n)
        g.step(t + 1, n)
        t += n
        a, b = util.full_state(o), util.full_state(g)
        util.assert_states_close(a, b, 1e-10, STATE_KEYS, f"bc edge step {t}")
        assert np.array_equal(a["disp"][top * 3 - 1], b["disp"][top * 3 - 1])          # prescribed values: bit-equal
    # past the table end the reference extrapolates SEGMENT 1 (time_index stays 1, J2:588-600): amp(t) = t / (40 dt)
    want = 0.02 * (t * dt) / (40 * dt)
    assert np.allclose(b["disp"][top * 3 - 1], want, rtol=1e-14)
    assert np.allclose(b["disp"][top[:4] * 3 - 3], -0.01 * (t * dt) / (40 * dt), rtol=1e-14)
    assert np.allclose(b["disp"][top[4:] * 3 - 3], 0.005 * (t * dt) / (40 * dt), rtol=1e-14)


# ---------------------------------------------------------------- checkpoint / resume (hakai_fem_b200/checkpoint.py)
def erosion_setup():
    model = ImpactDeck(plate=(8, 8, 2), proj=(3, 3, 3), v0=-900.0).build_model()
    model.MATERIAL[0].ductile = np.array([[0.02, 0.0, 30.0], [0.015, 0.4, 30.0]])
    return prepare(model)


def fracture_setup():
    return prepare(util.distorted_block(nx=5, ny=4, nz=6, jitter=0.05, ductile=True, strain_per_step=4e-4).build_model())


def check_resume(engine_cls, setup_fn, t_save, t_end, tmp_path, **prm):
    a = configure_engine(engine_cls, setup_fn(), **prm)
    a.step(1, t_save)
    n_before = len(a.deleted_ids())
    path = str(tmp_path / "ck.npz")
    _ck().save_checkpoint(a, path, t_save)
    a.step(t_save + 1, t_end - t_save)
    b = configure_engine(engine_cls, setup_fn(), **prm)
    t = _ck().load_checkpoint(b, path)
    assert t == t_save
    assert len(b.deleted_ids()) == n_before
    nb = b.step(t + 1, t_end - t)
    assert nb == len(a.deleted_ids()) - n_before, "only deletions after the checkpoint are reported as new"
    sa, sb = util.full_state(a), util.full_state(b)
    for k in STATE_KEYS:
        assert np.array_equal(np.asarray(sa[k]), np.asarray(sb[k])), k
    assert np.array_equal(a.deleted_ids(), b.deleted_ids())
    return a, b, n_before



def _ck():
    from hakai_fem_b200 import checkpoint
    return checkpoint


def case_checkpoint_resume(engine_cls, tmp_path):
    a, b, n_before = check_resume(engine_cls, erosion_setup, 57, 120, tmp_path)
    assert 0 < n_before < len(a.deleted_ids()), "checkpoint must fall between deletions"
    for c in range(2):
        pa, pb = a.contact_pair(c), b.contact_pair(c)
        for k in ("c_nodes_i", "c_nodes_j", "c_triangles", "c_triangles_eleid"):
            assert np.array_equal(pa[k], pb[k]), f"pair {c} {k}"
    a, _, n_before = check_resume(engine_cls, fracture_setup, 61, 100, tmp_path)
    assert 0 < n_before < len(a.deleted_ids())


def case_exact_mode_bitwise(engine_cls, cases=(("t5", 3000), ("crash_tube", 800), ("bullet_impact", 1500))):
    """hk_params.element_mode = 1 (reference-order element kernel, no FMA): the whole engine — nodal update, contact,
    element forces, triaxiality, deletion — is BIT-IDENTICAL to the oracle, including the self-contact deck whose hit
    set depends on the last bit of nodal positions."""
    for name, n in cases:
        st = prepare(util.t5_model()) if name == "t5" else util.deck_setup(name)
        o = configure_engine(OracleEngine, st)
        g = configure_engine(engine_cls, st, element_mode=1)
        assert o.step(1, n) == g.step(1, n)
        a, b = util.full_state(o), util.full_state(g)
        for k in a:
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), (name, k)
        assert np.array_equal(o.deleted_ids(), g.deleted_ids())
        assert o.counters()[1] == g.counters()[1]
