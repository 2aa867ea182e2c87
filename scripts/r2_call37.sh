#!/bin/bash
# the remaining SURVEY 8(d) lines on the final tree: I8 with friction, B1 plastic and elastic; + the new host-driver GPU test
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "host_driver_with_device_built_contact" 2>&1 | tail -3
timeout 1200 python bench.py --workload I8 --steps 40 --contact-myu 0.25 --no-cpu > gpurun_out/r2_bench_n1_I8_mu025.json 2> gpurun_out/r2_bench_n1_I8_mu025.err
timeout 600 python bench.py --workload B1 --steps 200 --warmup 50 --no-cpu > gpurun_out/r2_bench_n1_B1.json 2> gpurun_out/r2_bench_n1_B1.err
timeout 600 python bench.py --workload B1 --steps 200 --warmup 50 --strain-per-step 1e-6 --no-cpu > gpurun_out/r2_bench_n1_B1_elastic.json 2> gpurun_out/r2_bench_n1_B1_elastic.err
python - <<'PY'
import json
for w in ("I8_mu025","B1","B1_elastic"):
    try:
        j=json.loads(open(f"gpurun_out/r2_bench_n1_{w}.json").read().strip().splitlines()[-1])
    except Exception as ex:
        print(w, "FAILED", ex); continue
    r=j["roofline"]; c=j["config"]; e=j.get("e2e") or {}
    print(w, round(j["value"]/1e9,3),"G", round(j["ms_per_step"],4),"ms el",round(r["avg_launch_ms"],4),"frac",round(r["frac"],3),"step frac",round(r["whole_step"]["frac"],3),"nodal",round(r["nodal_kernel"]["ms_per_step"],4),c["regime"],"e2e",e.get("value"),(e.get("frame_loop") or {}).get("value"), j.get("contact",{}).get("ms_per_step") if j.get("contact") else None)
PY
for f in gpurun_out/r2_bench_n1_I8_mu025.err gpurun_out/r2_bench_n1_B1.err gpurun_out/r2_bench_n1_B1_elastic.err; do tail -n 2 $f; done
