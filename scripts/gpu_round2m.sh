#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02m_*
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $O/r02m_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02m_pytest_gpu.log
grep -E "^FAILED|passed|failed|rc=" $O/r02m_pytest_gpu.log | tail -20
timeout 120 build/igemm_trace 3 64 224 16 > $O/r02m_trace_conv1_1.txt 2>&1
sed -n 1,3p $O/r02m_trace_conv1_1.txt; sed -n 20,30p $O/r02m_trace_conv1_1.txt; tail -9 $O/r02m_trace_conv1_1.txt
run() { echo "== $MODEL $*" >> $O/r02m_lean.log; env "$@" timeout 300 python bench.py --lean --model $MODEL >> $O/r02m_lean.log 2>&1; }
for MODEL in vgg16 resnet50 mobilenet_v1; do run FCUDA_IGEMM_TMA_OUT=0; run FCUDA_IGEMM_TMA_OUT=1; done
grep -E "==|lean|Error|error|timed" $O/r02m_lean.log
