#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attn_gpu.py tests/test_attn_persist_ab_gpu.py tests/test_unet_gpu.py tests/test_decoder_layer_gpu.py tests/test_clip_splice_gpu.py -q -m gpu > gpurun_out/r03a_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r03a_tests.log
timeout 200 python scripts/bench_fa2.py > gpurun_out/r03a_attn_vs_flash_attn2.json 2> gpurun_out/r03a_fa2.err; cat gpurun_out/r03a_attn_vs_flash_attn2.json
/usr/bin/time -v timeout 900 python bench.py > gpurun_out/r03a_bench.json 2> gpurun_out/r03a_bench.err; echo "bench exit $?"; grep -E "Elapsed|bench +[0-9.]+s\]" gpurun_out/r03a_bench.err | tail -25
/usr/bin/time -v timeout 600 python bench.py --impl reference > gpurun_out/r03a_bench_ref.json 2> gpurun_out/r03a_bench_ref.err; echo "ref exit $?"; grep -E "Elapsed" gpurun_out/r03a_bench_ref.err; cat gpurun_out/r03a_bench_ref.json | cut -c1-600
