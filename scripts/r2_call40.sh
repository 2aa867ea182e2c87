#!/bin/bash
# contact kernels after the head preload / bbox grid change: contact parity tests, I8 bench, launch list
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "contact or reference_deck or erosion" 2>&1 | tail -3
timeout 1200 python bench.py --workload I8 --steps 40 --no-cpu > gpurun_out/r2_c40_I8.json 2> gpurun_out/r2_c40_I8.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2_c40_I8.json").read().strip().splitlines()[-1])
r=j["roofline"]
print("I8", round(j["value"]/1e9,3),"G", round(j["ms_per_step"],3),"ms el",round(r["avg_launch_ms"],3),"nodal",round(r["nodal_kernel"]["ms_per_step"],3),"contact",j["contact"]["ms_per_step"], j["contact"]["hits_per_step"], j["contact"]["tests_per_step"])
PY
CMD="python bench.py --workload I8 --steps 3 --warmup 3 --no-cpu --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 40 --csv --log-file gpurun_out/r2_launches_I8.csv $CMD > gpurun_out/r2_c40_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_I8.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
idx={h:i for i,h in enumerate(rows[hdr])}
for r in rows[hdr+1:][-18:]:
    try: v=float(r[idx["Metric Value"]])
    except: continue
    print(f"{r[idx['Kernel Name']].split('(')[0][:50]:50s} {v/1e3:9.1f} us {r[idx['Grid Size']]}")
PY
