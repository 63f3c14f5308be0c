#!/bin/bash
mkdir -p gpurun_out
DLLM_ATTN_NONPERSIST=1 timeout 300 python scripts/attn_ab_check.py save /tmp/attn_ref.pt 2>&1 | tail -1
for i in 1 2; do timeout 300 python scripts/attn_ab_check.py cmp /tmp/attn_ref.pt 2>&1 | tail -3 | cut -c1-300; done | tee gpurun_out/r02y_ab.log
timeout 400 python -m pytest tests/test_attn_gpu.py tests/test_unet_gpu.py tests/test_decoder_layer_gpu.py tests/test_clip_splice_gpu.py tests/test_sd_head_gpu.py -q -m gpu > gpurun_out/r02y_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r02y_tests.log
timeout 200 python scripts/bench_fa2.py > gpurun_out/r02y_attn_vs_flash_attn2.json 2> gpurun_out/r02y_fa2.err; cat gpurun_out/r02y_attn_vs_flash_attn2.json
