#!/bin/bash
mkdir -p gpurun_out
DLLM_ATTN_NONPERSIST=1 timeout 200 python scripts/attn_ab_check.py save /tmp/attn_ref.pt 2000 2>&1 | tail -1
DLLM_ATTN_DBG=0 timeout 200 python scripts/attn_ab_check.py cmp /tmp/attn_ref.pt 2000 2>&1 | tail -30 | cut -c1-400
