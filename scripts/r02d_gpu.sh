#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attn_gpu.py tests/test_gemm_gpu.py tests/test_decoder_layer_gpu.py -q -m gpu > gpurun_out/r02d_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r02d_tests.log
timeout 120 python scripts/bench_fa2.py > gpurun_out/r02d_fa2.json 2> gpurun_out/r02d_fa2.err; cat gpurun_out/r02d_fa2.json; tail -2 gpurun_out/r02d_fa2.err
timeout 120 python scripts/bench_gemm_bn.py > gpurun_out/r02d_bn_auto.json 2> gpurun_out/r02d_bn_auto.err; cat gpurun_out/r02d_bn_auto.json
DLLM_GEMM_BN=256 timeout 120 python scripts/bench_gemm_bn.py > gpurun_out/r02d_bn_256.json 2> gpurun_out/r02d_bn_256.err; cat gpurun_out/r02d_bn_256.json
DLLM_GEMM_BN=256 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02d_unet_launches.csv python scripts/unet_one_step.py > gpurun_out/r02d_unet_ncu.log 2>&1; echo "ncu exit $?"
