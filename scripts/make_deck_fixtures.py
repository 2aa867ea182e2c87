"""Builds tests/golden/deck_<name>.npz from the reference's own example decks (needs /root/reference).

Each fixture holds what `configure_engine` feeds through the C ABI for that deck — the arrays produced by
hakai_fem_b200.inp.read_inp_file + model_setup.prepare (the Python mirrors of readInpFile_j.jl and of hakai()'s
set-up) — so the GPU box, which has no /root/reference, can run the same decks.  tests/test_reference_decks.py
re-derives them from the decks when the reference tree is mounted and checks they are unchanged."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hakai_fem_b200.inp import read_inp_file                    # noqa: E402
from hakai_fem_b200.model_setup import prepare                  # noqa: E402

DECKS = {
    "bullet_impact": "HAKAI-v0.0.0/input/bullet-impact.inp",
    "metal_cutting": "HAKAI-v0.0.0/input/metal-cutting.inp",
    "charpy": "HAKAI-v0.0.1/input/Charpy-test-v0.0.1.inp",
    "projectile": "HAKAI-v0.0.1/input/projectile-impact-d1mm.inp",
    "car_crash_n2k": "HAKAI-v0.0.2/input/car-crash-N2k.inp",
    "crash_tube": "HAKAI-v0.0.1/input/crash-tube-80-350-solid.inp",
    "tensile_test": "HAKAI-v0.0.0/input/Tensile-test.inp",
}


def setup_to_arrays(st):
    m = st.model
    d = dict(coordmat=m.coordmat, elementmat=m.elementmat, element_material=m.element_material,
             element_instance=m.element_instance, diag_M=st.diag_M,
             scalars=np.array([st.d_time, st.time_num, st.elementMinSize, st.elementMaxSize, m.contact_flag,
                               len(m.MATERIAL), len(m.BC), len(m.IC), len(m.INSTANCE), len(st.CT)], float))
    for i, mt in enumerate(m.MATERIAL):
        d[f"mat{i}_s"] = np.array([mt.young, mt.poisson, mt.density])
        d[f"mat{i}_plastic"] = mt.plastic
        d[f"mat{i}_ductile"] = mt.ductile
    for i, bc in enumerate(m.BC):
        d[f"bc{i}_n"] = np.array([len(bc.dof), 1 if len(bc.amp_name) else 0])
        d[f"bc{i}_value"] = np.array(bc.value, float)
        d[f"bc{i}_amp"] = np.stack([np.asarray(bc.amplitude.time, float), np.asarray(bc.amplitude.value, float)])
        for j, dof in enumerate(bc.dof):
            d[f"bc{i}_dof{j}"] = np.asarray(dof, np.int64)
    for i, ic in enumerate(m.IC):
        d[f"ic{i}_value"] = np.array(ic.value, float)
        for j, dof in enumerate(ic.dof):
            d[f"ic{i}_dof{j}"] = np.asarray(dof, np.int64)
    if m.contact_flag >= 1:
        for i, ins in enumerate(m.INSTANCE):
            d[f"inst{i}_s"] = np.array([ins.node_offset, ins.nNode, ins.element_offset, ins.nElement, ins.material_id], np.int64)
            d[f"inst{i}_surfaces"] = ins.surfaces
            d[f"inst{i}_eleid"] = ins.surfaces_eleid
        for c, ct in enumerate(st.CT):
            d[f"ct{c}_s"] = np.array([ct.i_instance, ct.j_instance, ct.young], float)
            d[f"ct{c}_ni"] = ct.c_nodes_i
            d[f"ct{c}_nj"] = ct.c_nodes_j
            d[f"ct{c}_tri"] = ct.c_triangles
            d[f"ct{c}_te"] = ct.c_triangles_eleid
    return d


def main():
    out = os.path.join(ROOT, "tests", "golden")
    for name, rel in DECKS.items():
        model = read_inp_file(os.path.join("/root/reference", rel))
        cps = [(cp.instance_id_1, cp.instance_id_2, np.asarray(cp.elements_1, np.int64), np.asarray(cp.elements_2, np.int64))
               for cp in model.CP]                      # the deck's *Contact Pair list (empty: ALL EXTERIOR), before prepare()
        st = prepare(model)
        arrs = setup_to_arrays(st)
        arrs["cp_n"] = np.array([len(cps)], np.int64)   # what hk_build_contact (contact set-up on the device) is given
        for k, (i1, i2, e1, e2) in enumerate(cps):
            arrs[f"cp{k}_s"] = np.array([i1, i2], np.int64)
            arrs[f"cp{k}_e1"] = e1
            arrs[f"cp{k}_e2"] = e2
        np.savez_compressed(os.path.join(out, f"deck_{name}.npz"), **arrs)
        print(name, st.model.nNode, st.model.nElement, os.path.getsize(os.path.join(out, f"deck_{name}.npz")))


if __name__ == "__main__":
    main()
