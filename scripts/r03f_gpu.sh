#!/bin/bash
# r03f: final validation of the round: full GPU test tier, smoke(), attention microbench
mkdir -p gpurun_out
SECONDS=0
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r03f_gpu_tests.log 2>&1; echo "pytest exit $? after $SECONDS s"; tail -5 gpurun_out/r03f_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03f_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r03f_smoke.log
timeout 200 python scripts/bench_fa2.py > gpurun_out/r03f_attn_vs_flash_attn2.json 2> gpurun_out/r03f_fa2.err; cat gpurun_out/r03f_attn_vs_flash_attn2.json
