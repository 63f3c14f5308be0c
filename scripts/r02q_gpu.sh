#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_persist -s 3 -c 1 -o gpurun_out/r02q_fwd_persist python scripts/bench_attn.py > gpurun_out/r02q_ncu.log 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/r02q_fwd_persist.ncu-rep --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; d=rows[2]
for k in ('gpu__time_duration.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','launch__waves_per_multiprocessor','launch__registers_per_thread','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__grid_size','launch__shared_mem_per_block_dynamic','launch__shared_mem_per_block_static','launch__shared_mem_config_size','sm__maximum_warps_per_active_cycle_pct','sm__ctas_launched.sum'):
    if k in h: print(k, d[h.index(k)])
"
