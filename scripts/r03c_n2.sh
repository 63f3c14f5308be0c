#!/bin/bash
# r03c (2 GPUs): sharded optimizer / gradient reducer over NCCL, then the bench at N = 2
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_zero_nccl_gpu.py -q -m gpu -x > gpurun_out/r03c_zero_nccl.log 2>&1; echo "nccl test exit $?"; tail -4 gpurun_out/r03c_zero_nccl.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r03c_n2_stdout.log 2> gpurun_out/r03c_n2.err
echo "n2 exit $?"
grep -E '^\{' gpurun_out/r03c_n2_stdout.log | tail -1 > gpurun_out/r03c_bench_n2.json
grep "\[bench" gpurun_out/r03c_n2.err | tail -8
head -c 700 gpurun_out/r03c_bench_n2.json
