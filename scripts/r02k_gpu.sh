#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/ubench_mma scripts/ubench_mma_sm100.cu && timeout 60 /tmp/ubench_mma > gpurun_out/r02k_ubench_mma.json 2> gpurun_out/r02k_ubench_mma.err
echo "ubench exit $?"; cat gpurun_out/r02k_ubench_mma.json
