#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/r2_c9_multi.log
cat gpurun_out/r2_c9_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2_c9_bench_n2.json 2> gpurun_out/r2_c9_bench_n2.err
tail -c 3000 gpurun_out/r2_c9_bench_n2.json; grep -i "parity_check\|error\|Traceback" gpurun_out/r2_c9_bench_n2.err | head
