#!/bin/bash
# round 2ag: lanes of the filter-TMA warp (issue rate vs prefetch distance), interleaved repeats on one box
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02ag_*
timeout 200 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -p no:cacheprovider -k "sgeconv" 2>&1 | tail -1
run() { echo "== $MODEL $*" >> $O/r02ag_lean.log; env "$@" timeout 300 python bench.py --lean --model $MODEL >> $O/r02ag_lean.log 2>&1; }
for MODEL in vgg16 resnet50; do run FCUDA_IGEMM_TMA_LANES=4; run FCUDA_IGEMM_TMA_LANES=2; run FCUDA_IGEMM_TMA_LANES=1; run FCUDA_IGEMM_TMA_LANES=4; run FCUDA_IGEMM_TMA_LANES=2; done
grep -E "==|lean|rror|timed" $O/r02ag_lean.log
