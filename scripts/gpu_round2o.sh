#!/bin/bash
# round 2o: pointwise TMA slab + Winograd transform MLP variants: full parity, A/B lean bench, per-kernel launch lists
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02o_*
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > $O/r02o_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02o_pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" $O/r02o_pytest_gpu.log | tail -25
run() { echo "== $MODEL $*" >> $O/r02o_lean.log; env "$@" timeout 300 python bench.py --lean --model $MODEL >> $O/r02o_lean.log 2>&1; }
MODEL=vgg16; run FCUDA_WINO_MLP=0; run FCUDA_WINO_MLP=1
MODEL=resnet50; run FCUDA_IGEMM_PW=0; run FCUDA_IGEMM_PW=1
MODEL=mobilenet_v1; run FCUDA_IGEMM_PW=0; run FCUDA_IGEMM_PW=1
grep -E "==|lean|Error|error|timed" $O/r02o_lean.log
bash scripts/launches_gpu.sh r02o vgg16 resnet50 mobilenet_v1
