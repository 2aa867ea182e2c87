#!/bin/bash
# full GPU suite + smoke with step replay on
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/r2_c48_pytest.log
cat gpurun_out/r2_c48_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
