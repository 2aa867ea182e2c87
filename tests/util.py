"""Shared helpers of the test-suite: fixture models, random states, state comparison."""
import json
import os

import numpy as np

from hakai_fem_b200 import inp as I
from hakai_fem_b200.model_setup import prepare, configure_engine
from hakai_fem_b200.mesh import StretchDeck, ImpactDeck, steel

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_T5 = "/root/reference/HAKAI-v0.0.0/input/Tensile5e.inp"


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def t5_model() -> I.Model:
    """Tensile5e.inp rebuilt from the committed fixture (the GPU box has no /root/reference)."""
    j = load_json("tensile5e_model.json")
    coord = np.array(j["coordmat"], float)
    em = np.array(j["elementmat"], np.int64)
    mats = []
    for x in j["materials"]:
        m = I.Material(name=x["name"], density=x["density"], young=x["young"], poisson=x["poisson"])
        m.plastic = np.array(x["plastic"], float).reshape(-1, 2)
        m.Hd = np.array(x["Hd"], float)
        m.ductile = np.array(x["ductile"], float).reshape(-1, 3)
        mats.append(m)
    bcs = []
    for b in j["bc"]:
        bc = I.BC(amp_name=b["amp_name"], amplitude=I.Amplitude(name=b["amp_name"], time=np.array(b["amp_time"]),
                                                               value=np.array(b["amp_value"])))
        bc.dof = [np.array(d, np.int64) for d in b["dof"]]
        bc.value = list(b["value"])
        bcs.append(bc)
    ics = []
    for c in j["ic"]:
        ics.append(I.IC(type=c["type"], dof=[np.array(d, np.int64) for d in c["dof"]], value=list(c["value"])))
    nN, nE = j["nNode"], j["nElement"]
    part = I.Part(name="Part-1", nNode=nN, coordmat=coord, nElement=nE, elementmat=em,
                  material_name=j["part_material"][0], material_id=int(j["element_material"][0]))
    inst = I.Instance(name="Part-1-1", part_name="Part-1", part_id=1, material_id=part.material_id, nNode=nN,
                      nElement=nE, elements=np.arange(1, nE + 1))
    return I.Model(PART=[part], INSTANCE=[inst], NSET=[], ELSET=[], SURFACE=[], AMPLITUDE=[], MATERIAL=mats, BC=bcs,
                   IC=ics, CP=[], nNode=nN, coordmat=coord, nElement=nE, elementmat=em,
                   element_material=np.array(j["element_material"], np.int64),
                   element_instance=np.array(j["element_instance"], np.int64), d_time=j["d_time"],
                   end_time=j["end_time"], mass_scaling=j["mass_scaling"], contact_flag=j["contact_flag"])


def distorted_block(nx=4, ny=3, nz=5, jitter=0.15, ductile=False, seed=7, strain_per_step=2e-4):
    """Small stretch deck with every interior node displaced (tests the general, non-cuboid element)."""
    mat = steel("steel_Ductile", ductile=[[0.03, 0.0, 30.0], [0.02, 0.3, 30.0]]) if ductile else steel()
    return StretchDeck(nx, ny, nz, h=1.0, material=mat, jitter=jitter, seed=seed, strain_per_step=strain_per_step)


def random_state(setup, seed=0, plastic_fraction=0.5, scale_u=1e-3):
    """A physically plausible random loop state for single-step parity tests."""
    rng = np.random.default_rng(seed)
    m = setup.model
    fn, nip = 3 * m.nNode, 8 * m.nElement
    disp = rng.normal(0, scale_u, fn)
    disp_pre = disp - rng.normal(0, scale_u * 0.05, fn)
    velo = (disp - disp_pre) / setup.d_time
    stress = rng.normal(0, 300.0, (6, nip))
    strain = rng.normal(0, 1e-3, (6, nip))
    eps = np.where(rng.random(nip) < plastic_fraction, rng.random(nip) * 0.2, 0.0)
    yld = 755.0 + 500.0 * eps
    Q = rng.normal(0, 1.0, fn)
    return dict(disp=disp, disp_pre=disp_pre, velo=velo, Q=Q, integ_stress=stress, integ_strain=strain,
                integ_eq_plastic_strain=eps, integ_yield_stress=yld)


def full_state(eng):
    d = eng.download()
    d.update(eng.download_ex(fields=("disp_pre", "Q", "external_force", "position", "integ_yield_stress")))
    return d


def rel_err(a, b, floor=1e-300):
    """max|a-b| relative to the field's max magnitude (not smaller than `floor`)."""
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    s = max(np.abs(a).max(), np.abs(b).max(), floor)
    return float(np.abs(a - b).max() / s)


def assert_states_close(a, b, tol, keys=None, what="", floors=None):
    """floors: per-field magnitude below which a field is rounding noise (e.g. stress of a body in rigid motion)."""
    bad = []
    floors = floors or {}
    for k in (keys or a.keys()):
        if k == "element_flag":
            if not np.array_equal(a[k], b[k]):
                bad.append((k, "flags differ"))
            continue
        e = rel_err(a[k], b[k], floors.get(k, 1e-300))
        if not e <= tol:
            bad.append((k, e))
    assert not bad, f"{what}: fields beyond tol {tol}: {bad}"


def make_pair(setup, engine_cls, oracle_cls, **params):
    return configure_engine(oracle_cls, setup, **params), configure_engine(engine_cls, setup, **params)


# ---- reference example decks as fixtures (scripts/make_deck_fixtures.py) ------------------------------------
DECK_FILES = {
    "bullet_impact": "HAKAI-v0.0.0/input/bullet-impact.inp",
    "metal_cutting": "HAKAI-v0.0.0/input/metal-cutting.inp",
    "charpy": "HAKAI-v0.0.1/input/Charpy-test-v0.0.1.inp",
    "projectile": "HAKAI-v0.0.1/input/projectile-impact-d1mm.inp",
    "car_crash_n2k": "HAKAI-v0.0.2/input/car-crash-N2k.inp",
    "crash_tube": "HAKAI-v0.0.1/input/crash-tube-80-350-solid.inp",
    "tensile_test": "HAKAI-v0.0.0/input/Tensile-test.inp",
}


def deck_setup(name):
    """Setup of one of the reference's example decks rebuilt from tests/golden/deck_<name>.npz."""
    from hakai_fem_b200.model_setup import Setup, ContactTriangle
    z = np.load(os.path.join(GOLDEN, f"deck_{name}.npz"))
    sc = z["scalars"]
    d_time, time_num, emin, emax = (float(v) for v in sc[:4])
    cflag, n_mat, n_bc, n_ic, n_inst, n_ct = (int(v) for v in sc[4:10])
    mats = []
    for i in range(n_mat):
        y, p, rho = z[f"mat{i}_s"]
        m = I.Material(name=f"m{i}", density=float(rho), young=float(y), poisson=float(p))
        m.plastic = z[f"mat{i}_plastic"].reshape(-1, 2)
        if m.plastic.shape[0] > 1:
            m.Hd = (m.plastic[1:, 0] - m.plastic[:-1, 0]) / (m.plastic[1:, 1] - m.plastic[:-1, 1])
        m.ductile = z[f"mat{i}_ductile"].reshape(-1, 3)
        mats.append(m)
    bcs = []
    for i in range(n_bc):
        nl, has_amp = (int(v) for v in z[f"bc{i}_n"])
        amp = z[f"bc{i}_amp"]
        bc = I.BC(amp_name="Amp" if has_amp else "", amplitude=I.Amplitude(name="Amp", time=amp[0], value=amp[1]))
        bc.dof = [z[f"bc{i}_dof{j}"] for j in range(nl)]
        bc.value = list(z[f"bc{i}_value"])
        bcs.append(bc)
    ics = []
    for i in range(n_ic):
        vals = list(z[f"ic{i}_value"])
        ics.append(I.IC(type="VELOCITY", dof=[z[f"ic{i}_dof{j}"] for j in range(len(vals))], value=vals))
    coord, em = z["coordmat"], z["elementmat"]
    insts = []
    if cflag >= 1:
        for i in range(n_inst):
            no, nn, eo, ne, mid = (int(v) for v in z[f"inst{i}_s"])
            insts.append(I.Instance(name=f"i{i}", node_offset=no, nNode=nn, element_offset=eo, nElement=ne, material_id=mid,
                                    surfaces=z[f"inst{i}_surfaces"], surfaces_eleid=z[f"inst{i}_eleid"]))
    model = I.Model(PART=[], INSTANCE=insts, NSET=[], ELSET=[], SURFACE=[], AMPLITUDE=[], MATERIAL=mats, BC=bcs, IC=ics,
                    CP=[], nNode=coord.shape[1], coordmat=coord, nElement=em.shape[1], elementmat=em,
                    element_material=z["element_material"], element_instance=z["element_instance"], d_time=d_time,
                    end_time=d_time * time_num, mass_scaling=1.0, contact_flag=cflag)
    st = Setup(model, d_time, time_num, None, z["diag_M"], emin, emax)
    # the deck's own *Contact Pair list (empty: ALL EXTERIOR), for hk_build_contact
    model.CP = [I.CP(instance_id_1=int(z[f"cp{k}_s"][0]), instance_id_2=int(z[f"cp{k}_s"][1]), elements_1=z[f"cp{k}_e1"],
                     elements_2=z[f"cp{k}_e2"]) for k in range(int(z["cp_n"][0]))] if "cp_n" in z else []
    for c in range(n_ct):
        a, b, young = z[f"ct{c}_s"]
        st.CT.append(ContactTriangle(int(a), int(b), z[f"ct{c}_ni"], z[f"ct{c}_nj"], z[f"ct{c}_tri"], z[f"ct{c}_te"], float(young)))
    return st
