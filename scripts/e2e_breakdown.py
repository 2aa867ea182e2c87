"""Wall-clock breakdown of the end-to-end frame path (upload, steps, download, node_output) on one GPU."""
import sys
import time
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hakai_fem_b200.engine import Engine                      # noqa: E402
from hakai_fem_b200.mesh import B1                             # noqa: E402
from hakai_fem_b200.model_setup import prepare, configure_engine   # noqa: E402


def main():
    deck = B1()
    st = prepare(deck.build_model())
    eng = configure_engine(Engine, st)
    nN, nE = st.model.nNode, st.model.nElement
    eng.step(1, 20)
    d = eng.download()
    x = eng.download_ex(fields=("disp_pre", "Q", "integ_yield_stress"))
    pin = {k: torch.empty(s, dtype=torch.float64).pin_memory().numpy() for k, s in dict(
        node_stress=(6, nN), node_strain=(6, nN), node_eq_plastic_strain=(nN,), node_mises_stress=(nN,),
        node_triax_stress=(nN,), inc_num=(nN,)).items()}

    def t(label, fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        print(f"{label:32s} {1e3 * (time.perf_counter() - t0):9.2f} ms", flush=True)
        return r
    for rep in range(3):
        t("upload_state", lambda: eng.upload_state(disp=d["disp"], disp_pre=x["disp_pre"], velo=d["velo"], Q=x["Q"],
                                                   integ_stress=d["integ_stress"], integ_strain=d["integ_strain"],
                                                   integ_eq_plastic_strain=d["integ_eq_plastic_strain"],
                                                   integ_yield_stress=x["integ_yield_stress"]))
        t("10 steps", lambda: eng.step(21, 10))
        t("download disp/velo/flag", lambda: eng.download(fields=("disp", "velo", "element_flag")))
        t("node_output (pinned out)", lambda: eng.node_output(out=pin))
        t("node_output (fresh arrays)", lambda: eng.node_output())
        t("download full frame", lambda: eng.download())


if __name__ == "__main__":
    main()
