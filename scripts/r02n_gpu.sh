#!/bin/bash
mkdir -p gpurun_out
DLLM_LIB_PATH=$PWD/dreamllm_b200/libdreamllm_sm100_trace.so DLLM_NVCC_EXTRA=-DDLLM_ATTN_TRACE timeout 200 python scripts/attn_cta_log.py > gpurun_out/r02n_attn_cta_log.json 2> gpurun_out/r02n_attn_cta_log.err; echo "cta log exit $?"; cat gpurun_out/r02n_attn_cta_log.json | head -120; tail -3 gpurun_out/r02n_attn_cta_log.err
