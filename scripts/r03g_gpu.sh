#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attn_gpu.py tests/test_attn_persist_ab_gpu.py tests/test_decoder_layer_gpu.py tests/test_unet_gpu.py tests/test_causal_lm_gpu.py tests/test_clip_splice_gpu.py -q -m gpu > gpurun_out/r03g_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r03g_tests.log
timeout 200 python scripts/bench_fa2.py > gpurun_out/r03g_attn_vs_flash_attn2.json 2> gpurun_out/r03g_fa2.err; cat gpurun_out/r03g_attn_vs_flash_attn2.json
timeout 600 python bench.py > gpurun_out/r03g_bench.json 2> gpurun_out/r03g_bench.err; echo "bench exit $?"
python3 - <<EOF
import json
for l in open("gpurun_out/r03g_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print("c5", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "c2", d["c2"]["ms_per_step"], "c3", d["c3"]["ms_per_step"], "c4", d["c4"]["ms_total"], "c1", d["c1"]["gpu_ms"], "cpu", d["cpu_baseline"]["value"])
EOF
