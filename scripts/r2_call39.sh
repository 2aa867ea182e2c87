#!/bin/bash
# ncu --set full of the contact kernels of an I8 step
CMD="python bench.py --workload I8 --steps 3 --warmup 3 --no-cpu --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:'hk_contact_narrow|hk_contact_bbox' -s 40 -c 4 -o gpurun_out/r2_prof_contact $CMD > gpurun_out/r2_c39_ncu.log 2>&1
tail -2 gpurun_out/r2_c39_ncu.log | cut -c1-200
