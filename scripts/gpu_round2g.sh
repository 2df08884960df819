#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02g_*
timeout 120 build/gemm_selftest > $O/r02g_gemm_selftest.log 2>&1; echo "selftest rc=$?" >> $O/r02g_gemm_selftest.log
grep -E "FAIL|PASSED|rc=" $O/r02g_gemm_selftest.log | tail -8
FCUDA_IGEMM_SLAB=1 timeout 120 python scripts/debug_slab.py >> $O/r02g_debug.log 2>&1; echo "rc=$?" >> $O/r02g_debug.log
cat $O/r02g_debug.log
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $O/r02g_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02g_pytest_gpu.log
grep -E "^FAILED|passed|failed|rc=" $O/r02g_pytest_gpu.log | tail -40
for m in "vgg16 4" "resnet50 2"; do timeout 300 python scripts/determinism_probe.py $m >> $O/r02g_determinism.log 2>&1; done
cat $O/r02g_determinism.log
run() { echo "== $MODEL $*" >> $O/r02g_lean.log; env "$@" timeout 300 python bench.py --lean --model $MODEL >> $O/r02g_lean.log 2>&1; }
MODEL=vgg16
run FCUDA_IGEMM_SLAB=0 FCUDA_GEMM_TMA_STORE=0
run FCUDA_IGEMM_SLAB=0 FCUDA_GEMM_TMA_STORE=1
run FCUDA_IGEMM_SLAB=1 FCUDA_GEMM_TMA_STORE=1
run FCUDA_IGEMM_SLAB=1 FCUDA_GEMM_TMA_STORE=1 FCUDA_GEMM_CLUSTER=2
MODEL=resnet50
run FCUDA_IGEMM_SLAB=0 FCUDA_GEMM_TMA_STORE=0
run FCUDA_IGEMM_SLAB=1 FCUDA_GEMM_TMA_STORE=1
MODEL=mobilenet_v1
run FCUDA_IGEMM_SLAB=1
grep -E "==|lean|Error|error" $O/r02g_lean.log
FCUDA_IGEMM_SLAB=1 timeout 600 python bench.py --model vgg16 --no-cpu-baseline > $O/r02g_bench_vgg16.json 2>$O/r02g_bench_vgg16.err
