#!/bin/bash
# narrow-phase grid = one warp per possible triangle: contact parity, small-deck rates, I8 line
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "contact or reference_deck or erosion" 2>&1 | tail -2
python scripts/small_deck_rate.py bullet_impact metal_cutting car_crash_n2k
timeout 1200 python bench.py --workload I8 --steps 40 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('I8', round(j['value']/1e9,3), round(j['ms_per_step'],3), 'contact', round(j['contact']['ms_per_step'],4), j['contact']['hits_per_step'])"
