#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02j_*
FCUDA_IGEMM_CG=2 timeout 150 python scripts/debug_slab.py >> $O/r02j_debug.log 2>&1; echo "rc=$?" >> $O/r02j_debug.log
cat $O/r02j_debug.log
run() { echo "== $MODEL $*" >> $O/r02j_lean.log; env "$@" timeout 300 python bench.py --lean --model $MODEL >> $O/r02j_lean.log 2>&1; }
MODEL=vgg16
run FCUDA_IGEMM_CG=2
MODEL=resnet50
run FCUDA_IGEMM_CG=2
MODEL=mobilenet_v1
run FCUDA_IGEMM_CG=1 FCUDA_DW_VEC=0
run FCUDA_IGEMM_CG=1 FCUDA_DW_VEC=1
run FCUDA_IGEMM_CG=2
grep -E "==|lean|Error|error|timed" $O/r02j_lean.log
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_net.py -m gpu -q -k "dw or mobilenet or depthwise" -p no:cacheprovider 2>&1 | tail -5
