#!/bin/bash
# final tree: full GPU suite, smoke, the default bench line
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/r2_c51_pytest.log
cat gpurun_out/r2_c51_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2_bench_n1_W16.json 2> gpurun_out/r2_bench_n1_W16.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2_bench_n1_W16.json").read().strip().splitlines()[-1])
r=j["roofline"]; e=j["e2e"]
print("W16", round(j["value"]/1e9,3),"G", round(j["ms_per_step"],3),"ms el",round(r["avg_launch_ms"],3),"frac",round(r["frac"],3),"step",round(r["whole_step"]["frac"],3),"nodal",round(r["nodal_kernel"]["ms_per_step"],3),"e2e",round(e["value"]/1e9,3),round(e["frame_loop"]["value"]/1e9,3),j["cpu_baseline"]["value"], j["clocks"], j["gpu_launches"])
PY
tail -n 2 gpurun_out/r2_bench_n1_W16.err
