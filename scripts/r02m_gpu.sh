#!/bin/bash
mkdir -p gpurun_out
DLLM_LIB_PATH=$PWD/dreamllm_b200/libdreamllm_sm100_trace.so DLLM_NVCC_EXTRA=-DDLLM_ATTN_TRACE timeout 200 python scripts/attn_trace_bwd.py > gpurun_out/r02m_attn_bwd_trace.txt 2> gpurun_out/r02m_attn_bwd_trace.err; echo "trace exit $?"
timeout 200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:attn_ -s 8 -c 8 --csv --log-file gpurun_out/r02m_attn_kernels.csv python scripts/bench_attn.py > gpurun_out/r02m_ncu.log 2>&1; echo "ncu exit $?"; grep -v "^==" gpurun_out/r02m_attn_kernels.csv | cut -d, -f5,13- | head -30
