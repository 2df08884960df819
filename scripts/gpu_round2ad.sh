#!/bin/bash
# round 2ad: TensorGEMM epilogue with direct 32-byte stores (no smem staging) A/B, new flattened-box SGECONV cases
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02ad_*
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_variants.py -m gpu -q -x -p no:cacheprovider -k "sgeconv or tensor_gemm" 2>&1 | tail -2
run() { echo "== $MODEL $*" >> $O/r02ad_lean.log; env "$@" timeout 300 python bench.py --lean --model $MODEL >> $O/r02ad_lean.log 2>&1; }
MODEL=vgg16; run FCUDA_GEMM_TMA_STORE=1; run FCUDA_GEMM_TMA_STORE=2; run FCUDA_GEMM_TMA_STORE=1; run FCUDA_GEMM_TMA_STORE=2
grep -E "==|lean|rror|timed" $O/r02ad_lean.log
