#!/bin/bash
# r02t: persistent backward attention kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attn_gpu.py -q -m gpu -x > gpurun_out/r02t_tests_attn.log 2>&1; echo "attn tests exit $?"; tail -4 gpurun_out/r02t_tests_attn.log
timeout 400 python -m pytest tests/test_decoder_layer_gpu.py tests/test_clip_splice_gpu.py tests/test_unet_gpu.py tests/test_causal_lm_gpu.py -q -m gpu > gpurun_out/r02t_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r02t_tests.log
timeout 200 python scripts/bench_fa2.py > gpurun_out/r02t_attn_vs_flash_attn2.json 2> gpurun_out/r02t_fa2.err; echo "fa2 exit $?"; cat gpurun_out/r02t_attn_vs_flash_attn2.json
DLLM_ATTN_NONPERSIST=1 timeout 200 python scripts/bench_fa2.py > gpurun_out/r02t_attn_nonpersist.json 2> gpurun_out/r02t_fa2np.err; echo "fa2 nonpersist exit $?"; cat gpurun_out/r02t_attn_nonpersist.json
