#!/bin/bash
mkdir -p gpurun_out
for v in 13 25; do HK_ELEMENT_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "single_step or fracture_block or roundtrip or erosion or state_summary" 2>&1 | tail -2; done
python scripts/ab_element.py --configs "11,0,;12,0,;13,0,;20,0,;25,0," --rounds 2 --steps 30 > gpurun_out/r2_c12_ab.log 2>/dev/null; tail -6 gpurun_out/r2_c12_ab.log
timeout 900 python bench.py --workload F16D --steps 40 --no-cpu --no-e2e > gpurun_out/r2_c12_F16D.json 2>/dev/null
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2_c12_F16D.json").read().strip().splitlines()[-1])
r=j["roofline"]; c=j["config"]
print("F16D", round(j["value"]/1e9,3),"G", round(j["ms_per_step"],3),"ms el",round(r["avg_launch_ms"],3),"nodal",round(r["nodal_kernel"]["ms_per_step"],3),"del pass",round(r["deletion_pass_ms_per_step"],4),"live",c["live_elements_start"],c["live_elements_end"], c["deleted_per_step"][:6])
PY
