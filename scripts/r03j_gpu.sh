#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python -m pytest tests/test_gemm_gpu.py tests/test_attn_gpu.py -q -m gpu 2>&1 | tail -2
