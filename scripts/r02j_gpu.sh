#!/bin/bash
# r02j: forward-attention CTA timeline (trace build), LPT-order A/B via the attention microbench, robust C1 with/without split-K
mkdir -p gpurun_out
DLLM_LIB_PATH=$PWD/dreamllm_b200/libdreamllm_sm100_trace.so DLLM_NVCC_EXTRA=-DDLLM_ATTN_TRACE timeout 200 python scripts/attn_trace.py > gpurun_out/r02j_attn_trace.json 2> gpurun_out/r02j_attn_trace.txt; echo "trace exit $?"
timeout 200 python -m pytest tests/test_attn_gpu.py -q -m gpu > gpurun_out/r02j_tests.log 2>&1; echo "attn tests exit $?"; tail -3 gpurun_out/r02j_tests.log
timeout 200 python scripts/bench_fa2.py > gpurun_out/r02j_attn_vs_flash_attn2.json 2> gpurun_out/r02j_fa2.err; echo "fa2 exit $?"; cat gpurun_out/r02j_attn_vs_flash_attn2.json
timeout 200 python bench.py --only c1 --no-cpu-baseline > gpurun_out/r02j_c1_split.json 2> gpurun_out/r02j_c1_split.err; echo "c1 exit $?"
DLLM_GEMM_NO_SPLITK=1 timeout 200 python bench.py --only c1 --no-cpu-baseline > gpurun_out/r02j_c1_nosplit.json 2> gpurun_out/r02j_c1_nosplit.err; echo "c1 nosplit exit $?"
