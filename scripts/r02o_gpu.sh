#!/bin/bash
# r02o: persistent forward attention kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attn_gpu.py tests/test_kvcache_gpu.py tests/test_decoder_layer_gpu.py tests/test_clip_splice_gpu.py -q -m gpu -x > gpurun_out/r02o_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/r02o_tests.log
timeout 200 python scripts/bench_fa2.py > gpurun_out/r02o_attn_vs_flash_attn2.json 2> gpurun_out/r02o_fa2.err; echo "fa2 exit $?"; cat gpurun_out/r02o_attn_vs_flash_attn2.json
DLLM_ATTN_NONPERSIST=1 timeout 200 python scripts/bench_fa2.py > gpurun_out/r02o_attn_nonpersist.json 2> gpurun_out/r02o_fa2np.err; echo "fa2 nonpersist exit $?"; cat gpurun_out/r02o_attn_nonpersist.json
