"""Generates tests/golden/*.json from the reference's own input deck and the CPU oracle.

Run in the build container (needs /root/reference):  python scripts/make_golden.py
  tensile5e_model.json  — the arrays readInpFile produces for HAKAI-v0.0.0/input/Tensile5e.inp, as parsed
                          by hakai_fem_b200.inp (the Python mirror of readInpFile_j.jl); lets the GPU box,
                          which has no /root/reference, rebuild the same model.
  tensile5e_oracle.json — oracle snapshots (C++ restatement, triax by invariants) at a few steps.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hakai_fem_b200.inp import read_inp_file          # noqa: E402
from hakai_fem_b200.model_setup import prepare, configure_engine   # noqa: E402
from oracle.oracle_engine import OracleEngine          # noqa: E402

DECK = "/root/reference/HAKAI-v0.0.0/input/Tensile5e.inp"
OUT = os.path.join(ROOT, "tests", "golden")


def model_to_dict(m):
    return dict(
        nNode=m.nNode, nElement=m.nElement, coordmat=m.coordmat.tolist(), elementmat=m.elementmat.tolist(),
        element_material=m.element_material.tolist(), element_instance=m.element_instance.tolist(),
        d_time=m.d_time, end_time=m.end_time, mass_scaling=m.mass_scaling, contact_flag=m.contact_flag,
        materials=[dict(name=x.name, density=x.density, young=x.young, poisson=x.poisson, plastic=x.plastic.tolist(),
                        Hd=np.asarray(x.Hd).tolist(), ductile=x.ductile.tolist()) for x in m.MATERIAL],
        bc=[dict(amp_name=b.amp_name, amp_time=np.asarray(b.amplitude.time).tolist(),
                 amp_value=np.asarray(b.amplitude.value).tolist(), dof=[d.tolist() for d in b.dof], value=list(b.value))
            for b in m.BC],
        ic=[dict(type=i.type, dof=[d.tolist() for d in i.dof], value=list(i.value)) for i in m.IC],
        part_material=[p.material_name for p in m.PART],
    )


def main():
    os.makedirs(OUT, exist_ok=True)
    m = read_inp_file(DECK)
    with open(os.path.join(OUT, "tensile5e_model.json"), "w") as f:
        json.dump(model_to_dict(m), f)
    st = prepare(m)
    eng = configure_engine(OracleEngine, st)
    snaps = {}
    t = 0
    first_yield = None
    for target in (1, 2, 10, 316, 1000, 5000, 15152, 15153, 20000):
        while t < target:
            t += 1
            eng.step(t, 1)
            if first_yield is None and target <= 1000:
                if eng.download(fields=("integ_eq_plastic_strain",))["integ_eq_plastic_strain"].max() > 0:
                    first_yield = t
        d = eng.download()
        snaps[str(target)] = dict(disp=d["disp"].tolist(), eps=d["integ_eq_plastic_strain"].tolist(),
                                  stress=d["integ_stress"].T.reshape(-1).tolist(),
                                  triax=d["integ_triax_stress"].tolist(), element_flag=d["element_flag"].tolist())
    out = dict(first_yield_step=first_yield, deleted=eng.deleted_ids().tolist(), snapshots=snaps,
               elementVolume=st.elementVolume.tolist(), diag_M=st.diag_M.tolist(),
               elementMinSize=st.elementMinSize, elementMaxSize=st.elementMaxSize, time_num=st.time_num)
    with open(os.path.join(OUT, "tensile5e_oracle.json"), "w") as f:
        json.dump(out, f)
    print("first yield", first_yield, "deleted", out["deleted"])


if __name__ == "__main__":
    main()
