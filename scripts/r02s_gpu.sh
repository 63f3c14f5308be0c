#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attn_gpu.py tests/test_kvcache_gpu.py tests/test_decoder_layer_gpu.py tests/test_clip_splice_gpu.py tests/test_unet_gpu.py -q -m gpu -x > gpurun_out/r02s_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r02s_tests.log
for b in 2 8; do
ATTN_B=$b timeout 100 python scripts/bench_attn.py 2>&1 | head -1 | sed "s/^/B=$b persist: /"
ATTN_B=$b DLLM_ATTN_NONPERSIST=1 timeout 100 python scripts/bench_attn.py 2>&1 | head -1 | sed "s/^/B=$b nonpersist: /"
done
DLLM_LIB_PATH=$PWD/dreamllm_b200/libdreamllm_sm100_trace.so DLLM_NVCC_EXTRA=-DDLLM_ATTN_TRACE timeout 200 python scripts/attn_trace_persist.py > gpurun_out/r02s_persist_trace.txt 2> gpurun_out/r02s_persist_trace.err; echo "trace exit $?"; cat gpurun_out/r02s_persist_trace.txt; tail -3 gpurun_out/r02s_persist_trace.err
