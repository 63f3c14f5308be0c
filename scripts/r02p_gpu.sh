#!/bin/bash
mkdir -p gpurun_out
for sk in 0 300 550 800; do
DLLM_ATTN_SKEW=$sk timeout 100 python scripts/bench_attn.py 2>&1 | head -1 | sed "s/^/skew $sk: /"
done
DLLM_ATTN_NONPERSIST=1 timeout 100 python scripts/bench_attn.py 2>&1 | head -1 | sed "s/^/nonpersist: /"
