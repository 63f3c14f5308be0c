#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum
for mode in 0 1; do
DLLM_ATTN_NONPERSIST=$mode timeout 200 ncu --metrics $M --clock-control none -k regex:attn_ -s 8 -c 8 --csv --log-file gpurun_out/r02u_k$mode.csv python scripts/bench_attn_bwd_only.py > gpurun_out/r02u_ncu$mode.log 2>&1; echo "ncu exit $?"
python3 - <<EOF
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/r02u_k$mode.csv") if not l.startswith("=="))]
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value")
cur={}
for r in rows[1:]:
    cur.setdefault((r[0],r[ki][:40]),{})[r[mi].split('.')[0][:26]]=r[vi]
for k,v in cur.items(): print("nonpersist=$mode",k,v)
EOF
done
# UNet shapes: d=64 non-causal S=4096 (B=4 x 5 heads) and S=1024
for mode in 0 1; do
ATTN_B=4 ATTN_S=4096 ATTN_NH=5 ATTN_D=64 ATTN_CAUSAL=0 DLLM_ATTN_NONPERSIST=$mode timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_ -s 8 -c 4 --csv --log-file gpurun_out/r02u_unet$mode.csv python scripts/bench_attn_bwd_only.py > /dev/null 2>&1
grep -v "^==" gpurun_out/r02u_unet$mode.csv | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[1:]: print('unet d64 S4096 nonpersist=$mode', r[h.index('Kernel Name')][:36], r[h.index('Metric Value')])"
done
