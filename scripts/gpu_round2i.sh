#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02i_*
for cg in 1 2; do
  FCUDA_IGEMM_CG=$cg timeout 150 python scripts/debug_slab.py >> $O/r02i_debug.log 2>&1; echo "rc=$?" >> $O/r02i_debug.log
done
cat $O/r02i_debug.log
run() { echo "== $MODEL $*" >> $O/r02i_lean.log; env "$@" timeout 300 python bench.py --lean --model $MODEL >> $O/r02i_lean.log 2>&1; }
MODEL=vgg16
run FCUDA_IGEMM_CG=1
run FCUDA_IGEMM_CG=2
MODEL=resnet50
run FCUDA_IGEMM_CG=1
run FCUDA_IGEMM_CG=2
MODEL=mobilenet_v1
run FCUDA_IGEMM_CG=1
run FCUDA_IGEMM_CG=2
grep -E "==|lean|Error|error|timed" $O/r02i_lean.log
