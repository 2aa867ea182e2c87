#!/bin/bash
CMD="python bench.py --workload N128,128,256 --steps 4 --warmup 25 --no-cpu --no-e2e"
$CMD > gpurun_out/r2_c13_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'hk_element' -s 30 -c 1 -o gpurun_out/r2_prof_v13b $CMD > gpurun_out/r2_c13_ncu.log 2>&1
tail -2 gpurun_out/r2_c13_ncu.log
