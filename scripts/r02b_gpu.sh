#!/bin/bash
# one gpurun call: new attention kernels (TS) vs legacy, GN microbench, bench pieces with progress logs
mkdir -p gpurun_out
T="tests/test_attn_gpu.py tests/test_decoder_layer_gpu.py tests/test_causal_lm_gpu.py tests/test_unet_gpu.py tests/test_kvcache_gpu.py tests/test_clip_splice_gpu.py tests/test_sd_head_gpu.py"
timeout 500 python -m pytest $T -q -m gpu > gpurun_out/r02b_tests_ts.log 2>&1; echo "tests(TS) exit $?"; tail -6 gpurun_out/r02b_tests_ts.log
DLLM_ATTN_LEGACY=1 timeout 200 python -m pytest tests/test_attn_gpu.py -q -m gpu > gpurun_out/r02b_tests_legacy.log 2>&1; echo "tests(legacy) exit $?"; tail -2 gpurun_out/r02b_tests_legacy.log
timeout 120 python scripts/bench_fa2.py > gpurun_out/r02b_fa2_ts.json 2> gpurun_out/r02b_fa2_ts.err; cat gpurun_out/r02b_fa2_ts.json; tail -2 gpurun_out/r02b_fa2_ts.err
DLLM_ATTN_LEGACY=1 timeout 120 python scripts/bench_fa2.py > gpurun_out/r02b_fa2_legacy.json 2>/dev/null; cat gpurun_out/r02b_fa2_legacy.json
timeout 120 python scripts/bench_hbm_kernels.py > gpurun_out/r02b_hbm.json 2> gpurun_out/r02b_hbm.err
timeout 240 python bench.py --only c1,c5 --no-cpu-baseline > gpurun_out/r02b_bench_c5.json 2> gpurun_out/r02b_bench_c5.err; echo "bench c5 exit $?"; tail -4 gpurun_out/r02b_bench_c5.err
timeout 240 python bench.py --only c4 --no-cpu-baseline > gpurun_out/r02b_bench_c4.json 2> gpurun_out/r02b_bench_c4.err; echo "bench c4 exit $?"; tail -4 gpurun_out/r02b_bench_c4.err
timeout 240 python bench.py --only c2,c3 --no-cpu-baseline > gpurun_out/r02b_bench_c2.json 2> gpurun_out/r02b_bench_c2.err; echo "bench c2 exit $?"; tail -4 gpurun_out/r02b_bench_c2.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 --fast > gpurun_out/r02b_bench_ref.json 2> gpurun_out/r02b_bench_ref.err; echo "bench ref exit $?"; tail -6 gpurun_out/r02b_bench_ref.err
