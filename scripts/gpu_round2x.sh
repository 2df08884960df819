#!/bin/bash
# round 2x: next-barrier probes in the implicit GEMM's issuers / producers
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02x_*
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > $O/r02x_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02x_pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" $O/r02x_pytest_gpu.log | tail -25
run() { echo "== $MODEL $*" >> $O/r02x_lean.log; env "$@" timeout 300 python bench.py --lean --model $MODEL >> $O/r02x_lean.log 2>&1; }
for MODEL in vgg16 resnet50 mobilenet_v1; do run A=1; done
grep -E "==|lean|Error|error|timed" $O/r02x_lean.log
for pool in 0 1; do timeout 120 build/igemm_trace 64 64 224 16 3 $pool > $O/r02x_trace_conv1_2_bf16_pool$pool.txt 2>&1; sed -n 1,2p $O/r02x_trace_conv1_2_bf16_pool$pool.txt; sed -n 30,36p $O/r02x_trace_conv1_2_bf16_pool$pool.txt; tail -52 $O/r02x_trace_conv1_2_bf16_pool$pool.txt | head -24; tail -9 $O/r02x_trace_conv1_2_bf16_pool$pool.txt; done
timeout 120 build/igemm_trace 128 128 112 16 3 0 > $O/r02x_trace_conv2_2_bf16_pool0.txt 2>&1; sed -n 1,2p $O/r02x_trace_conv2_2_bf16_pool0.txt; tail -52 $O/r02x_trace_conv2_2_bf16_pool0.txt | head -24; tail -9 $O/r02x_trace_conv2_2_bf16_pool0.txt
