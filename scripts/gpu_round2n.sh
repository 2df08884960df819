#!/bin/bash
# round 2n: BF16x3 implicit GEMM — parity first (targeted), accuracy of the whole nets, then the lean bench in both modes
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02n_*
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_variants.py -m gpu -q --maxfail=10 -p no:cacheprovider -k "sgeconv or variants or forced or residual or fused_max_pool" > $O/r02n_pytest_conv.log 2>&1; echo "pytest rc=$?" >> $O/r02n_pytest_conv.log
grep -E "^FAILED|^ERROR|passed|failed|rc=|Error" $O/r02n_pytest_conv.log | tail -25
timeout 600 python scripts/accuracy_gpu.py > $O/r02n_accuracy.log 2>&1; cat $O/r02n_accuracy.log | tail -12
run() { echo "== $MODEL $*" >> $O/r02n_lean.log; timeout 300 python bench.py --lean --model $MODEL "$@" >> $O/r02n_lean.log 2>&1; }
for MODEL in vgg16 resnet50 mobilenet_v1; do run --precision tf32x3; run --precision fp32split; done
grep -E "==|lean|Error|error|timed" $O/r02n_lean.log
