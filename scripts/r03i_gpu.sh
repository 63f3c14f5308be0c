#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r03i_gpu_tests.log 2>&1; echo "pytest exit $? after $SECONDS s"; tail -3 gpurun_out/r03i_gpu_tests.log
timeout 400 python -m pytest tests/test_attn_gpu.py tests/test_kvcache_gpu.py tests/test_decoder_layer_gpu.py tests/test_unet_gpu.py tests/test_clip_splice_gpu.py -q -m gpu > gpurun_out/r03i_order2.log 2>&1; echo "order-2 exit $?"; tail -2 gpurun_out/r03i_order2.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
