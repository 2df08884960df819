#!/bin/bash
# round 2af: barrier-wait suspend hint A/B again, now that the TMA issue chain is gone (interleaved repeats, one box)
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02af_*
run() { echo "== $MODEL $*" >> $O/r02af_lean.log; env "$@" timeout 300 python bench.py --lean --model $MODEL >> $O/r02af_lean.log 2>&1; }
for MODEL in vgg16 resnet50; do run FCUDA_MBAR_SUSPEND_NS=100000; run FCUDA_MBAR_SUSPEND_NS=0; run FCUDA_MBAR_SUSPEND_NS=100000; run FCUDA_MBAR_SUSPEND_NS=0; run FCUDA_MBAR_SUSPEND_NS=500; done
grep -E "==|lean|rror|timed" $O/r02af_lean.log
