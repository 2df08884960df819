#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02l_*
timeout 600 python -m pytest tests/test_gpu_layers.py tests/test_gpu_conv.py -m gpu -q --maxfail=30 -p no:cacheprovider > $O/r02l_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02l_pytest_gpu.log
grep -E "^FAILED|passed|failed|rc=" $O/r02l_pytest_gpu.log | tail -20
timeout 120 build/igemm_trace 3 64 224 16 > $O/r02l_trace_conv1_1.txt 2>&1
head -50 $O/r02l_trace_conv1_1.txt
for m in vgg16 resnet50 mobilenet_v1; do timeout 300 python bench.py --lean --model $m >> $O/r02l_lean.log 2>&1; done
cat $O/r02l_lean.log
