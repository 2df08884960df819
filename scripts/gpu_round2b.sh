#!/bin/bash
# cluster-multicast TensorGEMM check + determinism probe
mkdir -p gpurun_out
O=gpurun_out
for cl in 0 1 2 4; do
  echo "== FCUDA_GEMM_CLUSTER=$cl" >> $O/r02b_gemm_selftest.log
  FCUDA_GEMM_CLUSTER=$cl timeout 120 build/gemm_selftest --bench >> $O/r02b_gemm_selftest.log 2>&1; echo "selftest rc=$?" >> $O/r02b_gemm_selftest.log
done
grep -E "==|rc=|FAIL|PASSED|GFLOP|TFLOP" $O/r02b_gemm_selftest.log | tail -60
for iss in 2 1; do
  echo "== FCUDA_IGEMM_ISSUERS=$iss" >> $O/r02b_determinism.log
  FCUDA_IGEMM_ISSUERS=$iss timeout 300 python scripts/determinism_probe.py vgg16 4 >> $O/r02b_determinism.log 2>&1
  FCUDA_IGEMM_ISSUERS=$iss timeout 300 python scripts/determinism_probe.py resnet50 2 >> $O/r02b_determinism.log 2>&1
done
FCUDA_GEMM_CLUSTER=1 timeout 300 python scripts/determinism_probe.py vgg16 4 nofusion >> $O/r02b_determinism.log 2>&1
cat $O/r02b_determinism.log
for cl in 1 2 4; do
  echo "== bench vgg16 lean FCUDA_GEMM_CLUSTER=$cl" >> $O/r02b_lean.log
  FCUDA_GEMM_CLUSTER=$cl timeout 300 python bench.py --model vgg16 --lean >> $O/r02b_lean.log 2>&1
done
FCUDA_GEMM_CLUSTER=2 timeout 300 python bench.py --model resnet50 --lean >> $O/r02b_lean.log 2>&1
FCUDA_GEMM_CLUSTER=1 timeout 300 python bench.py --model resnet50 --lean >> $O/r02b_lean.log 2>&1
cat $O/r02b_lean.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > $O/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02b_pytest_gpu.log
tail -8 $O/r02b_pytest_gpu.log
