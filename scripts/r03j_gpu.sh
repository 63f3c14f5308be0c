#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 150 python -m pytest tests/test_attn_gpu.py tests/test_kvcache_gpu.py -q -m gpu 2>&1 | tail -2
timeout 60 python scripts/bench_attn.py 2>&1 | tr '\n' ' '
