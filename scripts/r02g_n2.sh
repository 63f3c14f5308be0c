#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING,GRAPH
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02g_n2_stdout.log 2> gpurun_out/r02g_n2.err
echo "n2 exit $?"
grep -E '^\{' gpurun_out/r02g_n2_stdout.log | tail -1 > gpurun_out/r02g_bench_n2.json
grep -E "NCCL INFO" gpurun_out/r02g_n2_stdout.log gpurun_out/r02g_n2.err | grep -iE "algo|proto|nvls|channels|Ring|Tree|Connected|comm .* rank .* nranks|NCCL version|Using network|P2P|NVLS" | head -80 > gpurun_out/r02g_nccl_info.txt
grep "\[bench" gpurun_out/r02g_n2.err | tail -20
wc -l gpurun_out/r02g_nccl_info.txt; head -c 600 gpurun_out/r02g_bench_n2.json
unset NCCL_DEBUG NCCL_DEBUG_SUBSYS
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 -m pytest tests/test_zero_gpu.py -q -m gpu -x > gpurun_out/r02g_zero_n2.log 2>&1; tail -3 gpurun_out/r02g_zero_n2.log
