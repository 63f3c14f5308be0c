run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/n2_$tag.log 2>gpurun_out/n2_$tag.err; echo "$tag exit $?"; python -c "
import json
l=[x for x in open('gpurun_out/n2_$tag.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('$tag', round(d['ms_per_step'],1), round(d['value']), 'e2e', round(d['e2e']['value']), 'gemm', round(d['roofline']['achieved']))
else: print('$tag no result'); print(open('gpurun_out/n2_$tag.err').read()[-1500:])"; }
run D DLLM_NCCL_CTAS=0 DLLM_RESERVED_SMS=0
