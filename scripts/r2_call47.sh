#!/bin/bash
# step replay by CUDA graph: GPU parity suite with it on (default), small-deck rates on / off
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider 2>&1 | tail -4
echo "--- graph on"; python scripts/small_deck_rate.py
echo "--- graph off"; HK_STEP_GRAPH=0 python scripts/small_deck_rate.py
