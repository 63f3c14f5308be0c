#!/bin/bash
mkdir -p gpurun_out
DLLM_LIB_PATH=$PWD/dreamllm_b200/libdreamllm_sm100_trace.so DLLM_NVCC_EXTRA=-DDLLM_ATTN_TRACE timeout 200 python scripts/attn_trace_persist.py > gpurun_out/r02r_persist_trace.txt 2> gpurun_out/r02r_persist_trace.err; echo "trace exit $?"; cat gpurun_out/r02r_persist_trace.txt; tail -3 gpurun_out/r02r_persist_trace.err
for b in 2 3 8; do
ATTN_B=$b timeout 100 python scripts/bench_attn.py 2>&1 | head -1 | sed "s/^/B=$b persist: /"
ATTN_B=$b DLLM_ATTN_NONPERSIST=1 timeout 100 python scripts/bench_attn.py 2>&1 | head -1 | sed "s/^/B=$b nonpersist: /"
done
