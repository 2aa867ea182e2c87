#!/bin/bash
# N GPUs: bench.py under torchrun (parity_check first), final tree
N=${1:-2}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n${N}_W16.json 2> gpurun_out/r2_bench_n${N}_W16.err
python - <<PY
import json
j=json.loads(open("gpurun_out/r2_bench_n${N}_W16.json").read().strip().splitlines()[-1])
e=j["e2e"]
print(j["n_gpus"], round(j["value"]/1e9,3), "G", round(j["ms_per_step"],3), "ms frac", round(j["roofline"]["frac"],3), "e2e", round(e["value"]/1e9,3), e["seconds"], "frame_loop", round(e["frame_loop"]["value"]/1e9,3), j["config"]["cpu_affinity"], j["parity_check"]["ok"], j["clocks"])
print(j.get("per_rank_kernel_ms"))
PY
tail -n 3 gpurun_out/r2_bench_n${N}_W16.err
