#!/bin/bash
mkdir -p gpurun_out
T="tests/test_attn_gpu.py tests/test_unet_gpu.py tests/test_clip_splice_gpu.py tests/test_sd_head_gpu.py tests/test_kvcache_gpu.py tests/test_causal_lm_gpu.py"
timeout 500 python -m pytest $T -q -m gpu > gpurun_out/r02f_tests.log 2>&1; echo "tests exit $?"; grep -E "^FAILED|passed|failed" gpurun_out/r02f_tests.log | tail -8
timeout 240 python bench.py --only c4,c5 --no-cpu-baseline > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02f_bench.err
DLLM_STAGE1_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:dllm --csv --log-file gpurun_out/r02f_c5_launches.csv python bench.py --only c5 --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r02f_c5_ncu.log 2>&1; echo "ncu c5 exit $?"
timeout 240 compute-sanitizer --tool memcheck python scripts/sanitizer_targets.py > gpurun_out/r02f_memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/r02f_memcheck.log
timeout 400 compute-sanitizer --tool racecheck python scripts/sanitizer_targets.py > gpurun_out/r02f_racecheck.log 2>&1; echo "racecheck exit $?"; tail -3 gpurun_out/r02f_racecheck.log
timeout 300 ncu --set full --clock-control none -k "regex:gn_|sampler|attn_fwd|geglu|layernorm|upsample|copy_cols|gemm_kernel" -s 26 -c 26 -o gpurun_out/r02f_unet_full python scripts/ncu_targets_unet.py > gpurun_out/r02f_unet_ncu.log 2>&1; echo "ncu unet exit $?"
