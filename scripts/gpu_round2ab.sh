#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02ab_*
run() { echo "== $MODEL $*" >> $O/r02ab_lean.log; env "$@" timeout 200 python bench.py --lean --model $MODEL >> $O/r02ab_lean.log 2>&1; }
for MODEL in resnet50 mobilenet_v1; do run A=1; done
grep -E "==|lean|rror|timed" $O/r02ab_lean.log | head -20
timeout 600 python -m pytest tests -m gpu -q --maxfail=5 -p no:cacheprovider > $O/r02ab_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02ab_pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" $O/r02ab_pytest_gpu.log | tail -8
