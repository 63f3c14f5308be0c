#!/bin/bash
# r02i: micro-benchmarks for the attention softmax analysis + split-K fix validation + min-slices A/B
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/ubench_sm100 scripts/ubench_sm100.cu && timeout 120 gpurun_out/ubench_sm100 > gpurun_out/r02i_ubench.json 2> gpurun_out/r02i_ubench.err
echo "ubench exit $?"; cat gpurun_out/r02i_ubench.json | head -40
rm -f gpurun_out/ubench_sm100
timeout 500 python -m pytest tests/test_gemm_gpu.py tests/test_unet_gpu.py tests/test_sd_head_gpu.py -q -m gpu > gpurun_out/r02i_tests.log 2>&1
echo "tests exit $?"; tail -5 gpurun_out/r02i_tests.log
timeout 240 python bench.py --only c1,c4,c5 --no-cpu-baseline > gpurun_out/r02i_bench_min2.json 2> gpurun_out/r02i_bench_min2.err; echo "bench min2 exit $?"
DLLM_GEMM_SPLITK_MIN=3 timeout 240 python bench.py --only c1,c5 --no-cpu-baseline > gpurun_out/r02i_bench_min3.json 2> gpurun_out/r02i_bench_min3.err; echo "bench min3 exit $?"
DLLM_GEMM_SPLITK_MIN=4 timeout 240 python bench.py --only c1,c5 --no-cpu-baseline > gpurun_out/r02i_bench_min4.json 2> gpurun_out/r02i_bench_min4.err; echo "bench min4 exit $?"
