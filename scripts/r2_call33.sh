#!/bin/bash
# final bench lines (per-kernel CUDA events inside the timed region)
timeout 900 python bench.py > gpurun_out/r2_bench_n1_W16.json 2> gpurun_out/r2_bench_n1_W16.err
timeout 900 python bench.py --workload F16D --steps 40 --no-cpu > gpurun_out/r2_bench_n1_F16D.json 2> gpurun_out/r2_bench_n1_F16D.err
timeout 1200 python bench.py --workload I8 --steps 40 > gpurun_out/r2_bench_n1_I8.json 2> gpurun_out/r2_bench_n1_I8.err
python - <<'PY'
import json
for w in ("W16","F16D","I8"):
    try:
        j=json.loads(open(f"gpurun_out/r2_bench_n1_{w}.json").read().strip().splitlines()[-1])
    except Exception as ex:
        print(w, "FAILED", ex); continue
    r=j["roofline"]; c=j["config"]; e=j.get("e2e") or {}
    print(w, round(j["value"]/1e9,3),"G", round(j["ms_per_step"],3),"ms el",round(r["avg_launch_ms"],3),"frac",round(r["frac"],3),"step frac",round(r["whole_step"]["frac"],3),"nodal",round(r["nodal_kernel"]["ms_per_step"],3),"del",round(r["deletion_pass_ms_per_step"],3),"gap",round(r["launch_gaps_ms_per_step"],4),"e2e",e.get("value"),(e.get("frame_loop") or {}).get("value"), j.get("contact",{}).get("ms_per_step") if j.get("contact") else None, (j.get("cpu_baseline") or {}).get("value"), j["clocks"])
PY
for f in gpurun_out/r2_bench_n1_*.err; do tail -n 2 $f; done
