#!/bin/bash
# cluster-multicast TensorGEMM (fixed), slab producer, dual-accumulator issuers: correctness then speed
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/r02c_*
for cl in 0 2 4; do
  echo "== FCUDA_GEMM_CLUSTER=$cl" >> $O/r02c_gemm_selftest.log
  FCUDA_GEMM_CLUSTER=$cl timeout 120 build/gemm_selftest --bench >> $O/r02c_gemm_selftest.log 2>&1; echo "selftest rc=$?" >> $O/r02c_gemm_selftest.log
done
grep -E "==|rc=|FAIL|PASSED" $O/r02c_gemm_selftest.log | tail -30
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $O/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02c_pytest_gpu.log
grep -E "^FAILED|passed|failed|rc=" $O/r02c_pytest_gpu.log | tail -40
for m in "vgg16 4" "resnet50 2" "mobilenet_v1 2"; do
  timeout 300 python scripts/determinism_probe.py $m >> $O/r02c_determinism.log 2>&1
done
cat $O/r02c_determinism.log
run() { echo "== $*" >> $O/r02c_lean.log; env "$@" timeout 300 python bench.py --lean --model $MODEL >> $O/r02c_lean.log 2>&1; }
MODEL=vgg16
run FCUDA_IGEMM_SLAB=0 FCUDA_GEMM_CLUSTER=1
run FCUDA_IGEMM_SLAB=1 FCUDA_GEMM_CLUSTER=1
run FCUDA_IGEMM_SLAB=1 FCUDA_GEMM_CLUSTER=2
run FCUDA_IGEMM_SLAB=1 FCUDA_GEMM_CLUSTER=4
run FCUDA_IGEMM_SLAB=1 FCUDA_IGEMM_ISSUERS=1
run FCUDA_IGEMM_SLAB=1
MODEL=resnet50
run FCUDA_IGEMM_SLAB=0 FCUDA_GEMM_CLUSTER=1
run FCUDA_IGEMM_SLAB=1
MODEL=mobilenet_v1
run FCUDA_IGEMM_SLAB=1
grep -E "==|lean|Error|error" $O/r02c_lean.log
