#!/bin/bash
# 2 GPUs: all multi-GPU tests (incl. device-side erosion across ranks over the engine's NCCL communicator)
timeout 1500 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tee gpurun_out/r2_c31_multi.log | tail -15
