#!/bin/bash
# Run on the GPU box: build check + per-file pytest in separate processes with timeouts (a trap in one
# kernel must not take the rest of the run with it).  Output -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for f in "$@"; do
  name=$(basename $f .py)
  timeout 600 python -m pytest $f -x -q -m gpu > gpurun_out/$name.log 2>&1
  echo "$name exit $?" | tee -a gpurun_out/summary.txt
  tail -n 25 gpurun_out/$name.log
done
