#!/bin/bash
# round 2, GPU call 11: evidence set for the default build — full GPU suite, the three bench lines, ncu
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/r2_c11_pytest.log
cat gpurun_out/r2_c11_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_c11_smoke.log 2>&1; tail -2 gpurun_out/r2_c11_smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1_W16.json 2> gpurun_out/r2_bench_n1_W16.err
timeout 900 python bench.py --workload F16D --steps 40 --no-cpu > gpurun_out/r2_bench_n1_F16D.json 2> gpurun_out/r2_bench_n1_F16D.err
timeout 1200 python bench.py --workload I8 --steps 40 > gpurun_out/r2_bench_n1_I8.json 2> gpurun_out/r2_bench_n1_I8.err
timeout 600 python bench.py --impl reference --steps 4 --warmup 3 > gpurun_out/r2_bench_reference.json 2>/dev/null
CMD="python bench.py --workload N128,128,256 --steps 4 --warmup 25 --no-cpu --no-e2e"
$CMD > gpurun_out/r2_c11_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'hk_element|hk_nodal' -s 60 -c 2 -o gpurun_out/r2_prof_default $CMD > gpurun_out/r2_c11_ncu.log 2>&1
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD2 > gpurun_out/r2_c11_plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 52 -c 30 --csv --log-file gpurun_out/r2_launches_W16.csv $CMD2 > gpurun_out/r2_c11_ncu2.log 2>&1
python - <<'PY'
import json
for w in ("W16","F16D","I8"):
    try:
        j=json.loads(open(f"gpurun_out/r2_bench_n1_{w}.json").read().strip().splitlines()[-1])
    except Exception as ex:
        print(w, "FAILED", ex); continue
    r=j["roofline"]; c=j["config"]; e=j.get("e2e") or {}
    print(w, round(j["value"]/1e9,3),"G", round(j["ms_per_step"],3),"ms el",round(r["avg_launch_ms"],3),"frac",round(r["frac"],3),"step frac",round(r["whole_step"]["frac"],3),"nodal",round(r["nodal_kernel"]["ms_per_step"],3),c["regime"],c["untimed_steps_before_timing"],"live",c["live_elements_start"],c["live_elements_end"],"e2e",e.get("value"),e.get("seconds"),(e.get("frame_loop") or {}).get("value"), j.get("contact",{}).get("ms_per_step") if j.get("contact") else None, (j.get("cpu_baseline") or {}).get("value"))
PY
for f in gpurun_out/r2_bench_n1_*.err; do tail -n 2 $f; done
