#!/bin/bash
# final tree after the contact change: full GPU suite, smoke, I8 lines, I8 launch list
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/r2_c42_pytest.log
cat gpurun_out/r2_c42_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py --workload I8 --steps 40 > gpurun_out/r2_bench_n1_I8.json 2> gpurun_out/r2_bench_n1_I8.err
timeout 1200 python bench.py --workload I8 --steps 40 --contact-myu 0.25 --no-cpu > gpurun_out/r2_bench_n1_I8_mu025.json 2> gpurun_out/r2_bench_n1_I8_mu025.err
python - <<'PY'
import json
for w in ("I8","I8_mu025"):
    j=json.loads(open(f"gpurun_out/r2_bench_n1_{w}.json").read().strip().splitlines()[-1])
    r=j["roofline"]; e=j["e2e"]
    print(w, round(j["value"]/1e9,3),"G", round(j["ms_per_step"],3),"ms el",round(r["avg_launch_ms"],3),"frac",round(r["frac"],3),"step",round(r["whole_step"]["frac"],3),"nodal",round(r["nodal_kernel"]["ms_per_step"],3),"contact",round(j["contact"]["ms_per_step"],4), j["contact"]["hits_per_step"], j["contact"]["tests_per_step"],"e2e",round(e["value"]/1e9,3),round(e["frame_loop"]["value"]/1e9,3),(j.get("cpu_baseline") or {}).get("value"))
PY
CMD="python bench.py --workload I8 --steps 3 --warmup 3 --no-cpu --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 40 --csv --log-file gpurun_out/r2_launches_I8.csv $CMD > gpurun_out/r2_c42_ncu.log 2>&1
tail -1 gpurun_out/r2_c42_ncu.log | cut -c1-100
