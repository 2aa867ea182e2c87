#!/bin/bash
# round 2, GPU call 2: correctness of the ring-kernel variants on small parity cases, then W16 A/B
mkdir -p gpurun_out
for v in 21 11 20; do
  echo "== variant $v parity" >> gpurun_out/r2_c2_parity.log
  HK_ELEMENT_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider \
     -k "single_step or fracture_block or state_summary or roundtrip or contact_erosion" 2>&1 | tail -5 >> gpurun_out/r2_c2_parity.log
done
for v in 11 12 20 21 22 23; do
  echo "== variant $v" >> gpurun_out/r2_c2_bench.log
  HK_ELEMENT_VARIANT=$v timeout 400 python bench.py --steps 30 --warmup 20 --no-cpu --no-e2e >> gpurun_out/r2_c2_bench.log 2>> gpurun_out/r2_c2_bench.err
done
cat gpurun_out/r2_c2_parity.log
python - <<'PY'
import json
for l in open('gpurun_out/r2_c2_bench.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: j=json.loads(l)
    except Exception: continue
    r=j['roofline']; print(round(j['ms_per_step'],3), 'el', round(r['avg_launch_ms'],3), 'frac', round(r['frac'],3), 'nodal', round(r['nodal_kernel']['ms_per_step'],3), j['config']['regime'], j['config']['untimed_steps_before_timing'], j['clocks']['sm_mhz'])
PY
