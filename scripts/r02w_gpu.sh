#!/bin/bash
mkdir -p gpurun_out
DLLM_ATTN_NONPERSIST=1 timeout 200 python scripts/attn_ab_check.py save /tmp/attn_ref.pt 2>&1 | tail -2
timeout 200 python scripts/attn_ab_check.py cmp /tmp/attn_ref.pt > gpurun_out/r02w_ab.log 2>&1; echo "ab exit $?"; tail -25 gpurun_out/r02w_ab.log
for i in 1 2 3; do timeout 200 python -m pytest "tests/test_unet_gpu.py::test_unet_train_path_cond_gradient_vs_oracle_autograd" -q -m gpu 2>&1 | grep -E "dcond|passed|failed" | tr '\n' ' '; echo; done
echo "--- nonpersist"
for i in 1 2; do DLLM_ATTN_NONPERSIST=1 timeout 200 python -m pytest "tests/test_unet_gpu.py::test_unet_train_path_cond_gradient_vs_oracle_autograd" -q -m gpu -s 2>&1 | grep -E "dcond|passed|failed" | tr '\n' ' '; echo; done
