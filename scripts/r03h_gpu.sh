#!/bin/bash
# r03h: last GPU pass of the round — full GPU test tier on the final kernels, attention A/B timing, ncu --set full of the attention kernels
mkdir -p gpurun_out
SECONDS=0
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r03h_gpu_tests.log 2>&1; echo "pytest exit $? after $SECONDS s"; tail -3 gpurun_out/r03h_gpu_tests.log
for i in 1 2; do timeout 100 python scripts/bench_attn.py 2>&1 | tr '\n' ' '; echo; done
DLLM_ATTN_NONPERSIST=1 timeout 100 python scripts/bench_attn.py 2>&1 | tr '\n' ' '; echo " (one CTA per item)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 8 -c 4 -o gpurun_out/r03h_attn_final python scripts/bench_attn_bwd_only.py > gpurun_out/r03h_ncu.log 2>&1; echo "ncu exit $?"
