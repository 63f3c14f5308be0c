#!/bin/bash
# r02z: compute-sanitizer over the final kernels + ncu --set full of the persistent attention kernels
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck python scripts/sanitizer_targets.py > gpurun_out/r02z_memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/r02z_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python scripts/sanitizer_targets.py > gpurun_out/r02z_racecheck.log 2>&1; echo "racecheck exit $?"; tail -3 gpurun_out/r02z_racecheck.log
grep -c "Race reported\|hazard" gpurun_out/r02z_racecheck.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 8 -c 4 -o gpurun_out/r02z_attn_persist python scripts/bench_attn_bwd_only.py > gpurun_out/r02z_ncu.log 2>&1; echo "ncu exit $?"
