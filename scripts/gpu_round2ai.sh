#!/bin/bash
# round 2ai: CTA pairs (cta_group::2, 3xTF32 only) re-measured after the latency-chain work
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02ai_*
run() { echo "== $MODEL $*" >> $O/r02ai_lean.log; env "${@:2}" timeout 300 python bench.py --lean --model $MODEL --precision $1 >> $O/r02ai_lean.log 2>&1; }
for MODEL in vgg16 resnet50; do run tf32x3 FCUDA_IGEMM_CG=1; run tf32x3 FCUDA_IGEMM_CG=2; run tf32x3 FCUDA_IGEMM_CG=1; run tf32x3 FCUDA_IGEMM_CG=2; run fp32split A=1; done
grep -E "==|lean|rror|timed" $O/r02ai_lean.log
