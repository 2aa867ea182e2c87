#!/bin/bash
# round 2, GPU call 6: re-baseline after reverting the blocked layout, then ncu --set full of the default element kernel
mkdir -p gpurun_out
python scripts/ab_element.py --configs "11,0,;12,0,;20,0,;21,0," --rounds 2 --steps 30 > gpurun_out/r2_c6_ab.log 2> gpurun_out/r2_c6_ab.err
tail -6 gpurun_out/r2_c6_ab.log
CMD="python bench.py --workload N128,128,256 --steps 4 --warmup 25 --no-cpu --no-e2e"
$CMD > gpurun_out/r2_c6_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'hk_element|hk_nodal' -s 60 -c 2 -o gpurun_out/r2_prof_v20 $CMD > gpurun_out/r2_c6_ncu.log 2>&1
tail -3 gpurun_out/r2_c6_ncu.log
ls -la gpurun_out/*.ncu-rep
