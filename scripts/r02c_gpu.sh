#!/bin/bash
mkdir -p gpurun_out
T="tests/test_gemm_gpu.py tests/test_attn_gpu.py tests/test_decoder_layer_gpu.py tests/test_unet_gpu.py tests/test_sd_head_gpu.py tests/test_clip_splice_gpu.py tests/test_kvcache_gpu.py"
timeout 500 python -m pytest $T -q -m gpu > gpurun_out/r02c_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r02c_tests.log
timeout 120 python scripts/bench_fa2.py > gpurun_out/r02c_fa2.json 2> gpurun_out/r02c_fa2.err; cat gpurun_out/r02c_fa2.json; tail -2 gpurun_out/r02c_fa2.err
timeout 120 python scripts/bench_hbm_kernels.py > gpurun_out/r02c_hbm.json 2> gpurun_out/r02c_hbm.err
timeout 240 python bench.py --only c4,c5 --no-cpu-baseline > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r02c_bench.err
DLLM_STAGE1_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02c_c5_launches.csv python bench.py --only c5 --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r02c_c5_ncu.log 2>&1; echo "ncu exit $?"
