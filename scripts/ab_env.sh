#!/bin/bash
# A/B of environment settings on one box: scripts/ab_env.sh "VAR=a VAR=b ..." [workload]
WL=${2:-W16}
i=0
for kv in $1; do
  i=$((i+1))
  env ${kv//,/ } timeout 150 python bench.py --workload $WL --steps 15 --warmup 25 --no-cpu --no-e2e > gpurun_out/abe_$i.json 2> gpurun_out/abe_$i.err
  python -c "
import json;d=json.load(open('gpurun_out/abe_$i.json'));print('$kv: Gel/s %.3f step_ms %.3f elem_ms %.3f nodal_ms %.3f frac %.3f'%(d['value']/1e9,d['ms_per_step'],d['roofline']['avg_launch_ms'],d['roofline']['nodal_kernel']['avg_launch_ms'],d['roofline']['frac']))" || tail -3 gpurun_out/abe_$i.err
done
