#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_zero_nccl_gpu.py -q -m gpu -x > gpurun_out/r03e_zero_nccl.log 2>&1; echo "nccl test exit $?"; tail -6 gpurun_out/r03e_zero_nccl.log
