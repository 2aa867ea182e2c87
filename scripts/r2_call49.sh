#!/bin/bash
# where a step of a small contact deck goes: per-kernel durations (ncu launch list, plain launches)
HK_STEP_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 40 --csv --log-file gpurun_out/r2_launches_bullet.csv python scripts/small_deck_rate.py bullet_impact > gpurun_out/r2_c49.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_bullet.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
idx={h:i for i,h in enumerate(rows[hdr])}
tot=0
for r in rows[hdr+1:][:20]:
    try: v=float(r[idx["Metric Value"]])
    except: continue
    tot+=v
    print(f"{r[idx['Kernel Name']].split('(')[0][:50]:50s} {v/1e3:8.2f} us grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
print("sum of 20:", tot/1e3)
PY
