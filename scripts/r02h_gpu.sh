#!/bin/bash
mkdir -p gpurun_out
T="tests/test_gemm_gpu.py tests/test_unet_gpu.py tests/test_sd_head_gpu.py tests/test_decoder_layer_gpu.py tests/test_causal_lm_gpu.py tests/test_clip_splice_gpu.py"
timeout 500 python -m pytest $T -q -m gpu > gpurun_out/r02h_tests.log 2>&1; echo "tests exit $?"; grep -E "^FAILED|passed|failed" gpurun_out/r02h_tests.log | tail -8
timeout 240 python bench.py --only c1,c4,c5 --no-cpu-baseline > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02h_bench.err
DLLM_GEMM_NO_SPLITK=1 timeout 240 python bench.py --only c1,c5 --no-cpu-baseline > gpurun_out/r02h_bench_nosplit.json 2> gpurun_out/r02h_bench_nosplit.err; echo "bench(no split) exit $?"
timeout 200 ncu --set full --clock-control none -k "regex:gn_|layernorm_fwd|splitk" -c 14 -o gpurun_out/r02h_gn_full python scripts/ncu_targets_unet.py > gpurun_out/r02h_gn_ncu.log 2>&1; echo "ncu gn exit $?"
