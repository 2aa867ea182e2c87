#!/bin/bash
mkdir -p gpurun_out
HK_ELEMENT_VARIANT=25 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "single_step or fracture_block or roundtrip" 2>&1 | tail -3
HK_ELEMENT_VARIANT=27 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "single_step or fracture_block or roundtrip" 2>&1 | tail -3
HK_ELEMENT_VARIANT=13 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "single_step or fracture_block or roundtrip" 2>&1 | tail -3
python scripts/ab_element.py --configs "12,0,;13,0,;20,0,;24,0,;25,0,;26,0,;27,0," --rounds 2 --steps 30 > gpurun_out/r2_c7_ab.log 2> gpurun_out/r2_c7_ab.err
tail -9 gpurun_out/r2_c7_ab.log; tail -3 gpurun_out/r2_c7_ab.err
