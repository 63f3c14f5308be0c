#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 900 python bench.py > gpurun_out/r03b_bench.json 2> gpurun_out/r03b_bench.err; echo "bench exit $? after $SECONDS s"; grep -E "bench +[0-9.]+s\]" gpurun_out/r03b_bench.err | tail -30
SECONDS=0
timeout 600 python bench.py --impl reference > gpurun_out/r03b_bench_ref.json 2> gpurun_out/r03b_bench_ref.err; echo "ref exit $? after $SECONDS s"; cat gpurun_out/r03b_bench_ref.json | cut -c1-700
DLLM_STAGE1_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:dllm --csv --log-file gpurun_out/r03b_c5_launches.csv python bench.py --only c5 --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r03b_c5_ncu.log 2>&1; echo "ncu c5 exit $?"
