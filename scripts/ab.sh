#!/bin/bash
# A/B of element-kernel variants on one box: scripts/ab.sh "0 1 2 3 4 simple" [workload]
WL=${2:-W16}
for v in $1; do
  if [ "$v" = "simple" ]; then export HK_ELEMENT_KERNEL=simple; unset HK_ELEMENT_VARIANT; else unset HK_ELEMENT_KERNEL; export HK_ELEMENT_VARIANT=$v; fi
  timeout 150 python bench.py --workload $WL --steps 15 --warmup 25 --no-cpu --no-e2e > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python -c "
import json;d=json.load(open('gpurun_out/ab_$v.json'));print('variant $v: Gel/s %.3f step_ms %.3f elem_ms %.3f nodal_ms %.3f frac %.3f'%(d['value']/1e9,d['ms_per_step'],d['roofline']['avg_launch_ms'],d['roofline']['nodal_kernel']['avg_launch_ms'],d['roofline']['frac']),d['clocks'])" || tail -3 gpurun_out/ab_$v.err
done
