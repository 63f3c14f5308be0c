#!/bin/bash
# Round-2 A/B of the HBM-kernel build flags, one gpurun call:
#   gpurun --timeout 600 -- 'bash scripts/ab_hbm_flags.sh'
# 1) microbench with the shipped library; 2) build /tmp/ab.so with -DDLLM_VEC128 -DDLLM_GN_GROUP2; 3) run the parity tests that cover the
# touched kernels against it; 4) microbench again.  Results -> gpurun_out/ab_hbm_*.json / .log
mkdir -p gpurun_out
python scripts/bench_hbm_kernels.py > gpurun_out/ab_hbm_shipped.json 2> gpurun_out/ab_hbm_shipped.err
export DLLM_NVCC_EXTRA="-DDLLM_VEC128 -DDLLM_GN_GROUP2" DLLM_LIB_PATH=/tmp/ab.so
python -c "from dreamllm_b200 import _lib; print(_lib.build(force=True))" > gpurun_out/ab_hbm_build.log 2>&1 || { tail -5 gpurun_out/ab_hbm_build.log; exit 1; }
timeout 400 python -m pytest tests/test_elementwise_gpu.py tests/test_decoder_layer_gpu.py tests/test_unet_gpu.py tests/test_clip_splice_gpu.py \
  tests/test_sd_head_gpu.py -q -m gpu -x > gpurun_out/ab_hbm_tests.log 2>&1
echo "parity on the flagged build: exit $?"; tail -3 gpurun_out/ab_hbm_tests.log
python scripts/bench_hbm_kernels.py > gpurun_out/ab_hbm_flagged.json 2> gpurun_out/ab_hbm_flagged.err
python - <<'PY'
import json
a = json.load(open("gpurun_out/ab_hbm_shipped.json"))["kernels"]; b = json.load(open("gpurun_out/ab_hbm_flagged.json"))["kernels"]
for k in a:
    print(f"{k:38s} {a[k]['us']:9.1f} us {a[k]['frac']:.3f}  ->  {b[k]['us']:9.1f} us {b[k]['frac']:.3f}   x{a[k]['us'] / b[k]['us']:.2f}")
PY
