#!/bin/bash
# round 2, GPU call 3: full GPU suite (device-side erosion, ring kernel default), layout A/B, F16D and I8 bench lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2_c3_pytest.log
run() { echo "== $1" >> gpurun_out/r2_c3_bench.log; shift; env "$@" >> gpurun_out/r2_c3_bench.log 2>> gpurun_out/r2_c3_bench.err; }
run "v20 blocked" HK_ELEMENT_VARIANT=20 timeout 400 python bench.py --steps 30 --warmup 20 --no-cpu --no-e2e
run "v20 soa" HK_ELEMENT_VARIANT=20 HK_LAYOUT_BLOCKED=0 timeout 400 python bench.py --steps 30 --warmup 20 --no-cpu --no-e2e
run "v12 blocked" HK_ELEMENT_VARIANT=12 timeout 400 python bench.py --steps 30 --warmup 20 --no-cpu --no-e2e
run "v11 soa" HK_ELEMENT_VARIANT=11 HK_LAYOUT_BLOCKED=0 timeout 400 python bench.py --steps 30 --warmup 20 --no-cpu --no-e2e
run "F16D" timeout 600 python bench.py --workload F16D --steps 40 --warmup 20 --no-cpu --no-e2e
run "I8" timeout 900 python bench.py --workload I8 --steps 40 --warmup 12 --no-cpu --no-e2e
cat gpurun_out/r2_c3_pytest.log
python - <<'PY'
import json
for l in open('gpurun_out/r2_c3_bench.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: j=json.loads(l)
    except Exception: continue
    r=j['roofline']; c=j['config']
    print(' ', round(j['value']/1e9,3),'G', round(j['ms_per_step'],3), 'el', round(r['avg_launch_ms'],3), 'frac', round(r['frac'],3), 'nodal', round(r['nodal_kernel']['ms_per_step'],3), c['regime'], c['untimed_steps_before_timing'], c['live_elements_start'], c['live_elements_end'], j['clocks']['sm_mhz'], j.get('contact'))
    if 'deleted_per_step' in c: print('  deleted/step', c['deleted_per_step'])
PY
tail -3 gpurun_out/r2_c3_bench.err
