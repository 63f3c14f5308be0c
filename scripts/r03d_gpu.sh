#!/bin/bash
# r03d: one-launch (cooperative) GroupNorm forward
mkdir -p gpurun_out
DLLM_GN_NO_FUSE=1 timeout 200 python scripts/gn_fused_ab.py save /tmp/gn_ref.pt 2>&1 | tail -12
timeout 200 python scripts/gn_fused_ab.py cmp /tmp/gn_ref.pt 2>&1 | tail -13
timeout 500 python -m pytest tests/test_unet_gpu.py tests/test_sd_head_gpu.py -q -m gpu > gpurun_out/r03d_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r03d_tests.log
timeout 300 python bench.py --only c4,c5 --no-cpu-baseline > gpurun_out/r03d_bench.json 2> gpurun_out/r03d_bench.err; echo "bench exit $?"
DLLM_GN_NO_FUSE=1 timeout 300 python bench.py --only c4,c5 --no-cpu-baseline > gpurun_out/r03d_bench_nofuse.json 2> gpurun_out/r03d_bench_nofuse.err; echo "bench nofuse exit $?"
python3 - <<EOF
import json
for f in ("r03d_bench.json","r03d_bench_nofuse.json"):
    for l in open("gpurun_out/"+f):
        if l.startswith("{"):
            d=json.loads(l); print(f, "c5", d["ms_per_step"], "c4", d["c4"]["ms_total"])
EOF
