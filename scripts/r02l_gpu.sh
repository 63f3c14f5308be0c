#!/bin/bash
# r02l: single-thread (elect.sync) MMA / TMA issue + hoisted descriptors in the attention and GEMM kernels
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_attn_gpu.py tests/test_gemm_gpu.py tests/test_decoder_layer_gpu.py tests/test_kvcache_gpu.py -q -m gpu > gpurun_out/r02l_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r02l_tests.log
timeout 200 python scripts/bench_fa2.py > gpurun_out/r02l_attn_vs_flash_attn2.json 2> gpurun_out/r02l_fa2.err; echo "fa2 exit $?"; cat gpurun_out/r02l_attn_vs_flash_attn2.json
DLLM_LIB_PATH=$PWD/dreamllm_b200/libdreamllm_sm100_trace.so DLLM_NVCC_EXTRA=-DDLLM_ATTN_TRACE timeout 200 python scripts/attn_trace.py > gpurun_out/r02l_attn_trace.json 2> gpurun_out/r02l_attn_trace.txt; echo "trace exit $?"
timeout 120 python scripts/bench_gemm.py > gpurun_out/r02l_gemm.json 2> gpurun_out/r02l_gemm.err; echo "gemm exit $?"; tail -5 gpurun_out/r02l_gemm.json
timeout 120 python scripts/bench_gemm_bn.py > gpurun_out/r02l_bn_auto.json 2> gpurun_out/r02l_bn_auto.err; cat gpurun_out/r02l_bn_auto.json
DLLM_GEMM_BN=1 timeout 120 python scripts/bench_gemm_bn.py > gpurun_out/r02l_bn_minpad.json 2> gpurun_out/r02l_bn_minpad.err; cat gpurun_out/r02l_bn_minpad.json
timeout 300 python bench.py --only c1,c4,c5 --no-cpu-baseline > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "bench exit $?"
DLLM_GEMM_BN=1 timeout 300 python bench.py --only c4,c5 --no-cpu-baseline > gpurun_out/r02l_bench_minpad.json 2> gpurun_out/r02l_bench_minpad.err; echo "bench minpad exit $?"
