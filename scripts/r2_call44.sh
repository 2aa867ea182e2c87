#!/bin/bash
# deletion pass after the flush / count changes: fracture parity tests, F16D bench
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "fracture or erosion or t5 or reference_deck or bitwise" 2>&1 | tail -2
timeout 900 python bench.py --workload F16D --steps 40 --no-cpu > gpurun_out/r2_bench_n1_F16D.json 2> gpurun_out/r2_bench_n1_F16D.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2_bench_n1_F16D.json").read().strip().splitlines()[-1])
r=j["roofline"]; e=j["e2e"]
print("F16D", round(j["value"]/1e9,3),"G", round(j["ms_per_step"],3),"ms el",round(r["avg_launch_ms"],3),"nodal",round(r["nodal_kernel"]["ms_per_step"],3),"del",round(r["deletion_pass_ms_per_step"],3), j["config"]["live_elements_end"], "e2e", round(e["value"]/1e9,3), round(e["frame_loop"]["value"]/1e9,3))
PY
tail -n 2 gpurun_out/r2_bench_n1_F16D.err
