#!/bin/bash
# round 2aa: lockstep TMA issue of the pointwise slabs (2 k-blocks x 4 boxes per round), next-slot probe in the generic producers
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02aa_*
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > $O/r02aa_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02aa_pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" $O/r02aa_pytest_gpu.log | tail -25
run() { echo "== $MODEL $*" >> $O/r02aa_lean.log; env "$@" timeout 300 python bench.py --lean --model $MODEL >> $O/r02aa_lean.log 2>&1; }
for MODEL in resnet50 mobilenet_v1 vgg16; do run A=1; done
grep -E "==|lean|Error|error|timed" $O/r02aa_lean.log
