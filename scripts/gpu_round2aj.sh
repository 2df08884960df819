#!/bin/bash
# round 2aj: "plain" epilogue path (no residual, single accumulator): parity + lean bench
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02aj_*
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3
run() { echo "== $MODEL $*" >> $O/r02aj_lean.log; env "$@" timeout 300 python bench.py --lean --model $MODEL >> $O/r02aj_lean.log 2>&1; }
for MODEL in mobilenet_v1 resnet50 vgg16; do run A=1; done
grep -E "==|lean|rror|timed" $O/r02aj_lean.log
