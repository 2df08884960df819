#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02e_*
for s in 1; do for i in 2; do
  FCUDA_IGEMM_SLAB=$s FCUDA_IGEMM_ISSUERS=$i timeout 120 python scripts/debug_slab.py >> $O/r02e_debug.log 2>&1; echo "rc=$?" >> $O/r02e_debug.log
done; done
cat $O/r02e_debug.log
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $O/r02e_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02e_pytest_gpu.log
grep -E "^FAILED|passed|failed|rc=" $O/r02e_pytest_gpu.log | tail -40
for m in "vgg16 4" "resnet50 2" "mobilenet_v1 2"; do
  timeout 300 python scripts/determinism_probe.py $m >> $O/r02e_determinism.log 2>&1
done
cat $O/r02e_determinism.log
run() { echo "== $MODEL $*" >> $O/r02e_lean.log; env "$@" timeout 300 python bench.py --lean --model $MODEL >> $O/r02e_lean.log 2>&1; }
MODEL=vgg16
run FCUDA_IGEMM_SLAB=0 FCUDA_GEMM_CLUSTER=1
run FCUDA_IGEMM_SLAB=1 FCUDA_GEMM_CLUSTER=1
run FCUDA_IGEMM_SLAB=1 FCUDA_GEMM_CLUSTER=2
run FCUDA_IGEMM_SLAB=1 FCUDA_GEMM_CLUSTER=4
run FCUDA_IGEMM_SLAB=1
MODEL=resnet50
run FCUDA_IGEMM_SLAB=0 FCUDA_GEMM_CLUSTER=1
run FCUDA_IGEMM_SLAB=1
MODEL=mobilenet_v1
run FCUDA_IGEMM_SLAB=1
grep -E "==|lean|Error|error" $O/r02e_lean.log
