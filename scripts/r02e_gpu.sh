#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_unet_gpu.py tests/test_sd_head_gpu.py tests/test_gemm_gpu.py -q -m gpu > gpurun_out/r02e_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r02e_tests.log
timeout 240 python bench.py --only c4,c5 --no-cpu-baseline > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02e_bench.err
DLLM_STAGE1_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dllm --csv --log-file gpurun_out/r02e_c5_launches.csv python bench.py --only c5 --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r02e_c5_ncu.log 2>&1; echo "ncu c5 exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 4 -c 4 -o gpurun_out/r02e_attn_ts python scripts/ncu_targets.py > gpurun_out/r02e_ncu.log 2>&1; echo "ncu attn exit $?"
