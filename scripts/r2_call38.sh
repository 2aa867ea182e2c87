#!/bin/bash
# launch list of an I8 run (contact kernels one by one)
CMD="python bench.py --workload I8 --steps 3 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/r2_c38_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 80 --csv --log-file gpurun_out/r2_launches_I8.csv $CMD > gpurun_out/r2_c38_ncu.log 2>&1
tail -1 gpurun_out/r2_c38_plain.log | cut -c1-300
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_I8.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
idx={h:i for i,h in enumerate(rows[hdr])}
seq=[]
for r in rows[hdr+1:]:
    try: v=float(r[idx["Metric Value"]])
    except: continue
    seq.append((r[idx["Kernel Name"]].split("(")[0][:50], v, r[idx["Grid Size"]] if "Grid Size" in idx else ""))
for k,v,g in seq[-44:]: print(f"{k:50s} {v/1e3:9.1f} us {g}")
PY
