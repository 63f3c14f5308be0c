#!/bin/bash
# r02v: stale-max fast path in the persistent forward kernel, vectorised backward prep kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attn_gpu.py tests/test_kvcache_gpu.py tests/test_decoder_layer_gpu.py tests/test_unet_gpu.py tests/test_clip_splice_gpu.py -q -m gpu > gpurun_out/r02v_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r02v_tests.log
timeout 200 python scripts/bench_fa2.py > gpurun_out/r02v_attn_vs_flash_attn2.json 2> gpurun_out/r02v_fa2.err; echo "fa2 exit $?"; cat gpurun_out/r02v_attn_vs_flash_attn2.json
M=gpu__time_duration.sum
timeout 200 ncu --metrics $M --clock-control none -k regex:attn_ -s 8 -c 4 --csv --log-file gpurun_out/r02v_k.csv python scripts/bench_attn_bwd_only.py > gpurun_out/r02v_ncu.log 2>&1
grep -v "^==" gpurun_out/r02v_k.csv | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[1:]: print(r[h.index('Kernel Name')][:40], r[h.index('Metric Value')])"
timeout 300 python bench.py --only c1,c4,c5 --no-cpu-baseline > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; echo "bench exit $?"
python3 -c "
import json
for l in open('gpurun_out/r02v_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print('c5', d['ms_per_step'], 'c4', d['c4']['ms_total'], 'c1', d['c1']['gpu_ms'], d['c1']['gpu_ms_eager'])
"
