#!/bin/bash
# round 2ac: warp-uniform issuer loop — conv2_2 / conv1_2 traces (compare with r02x/r02y: 0.185 ms / 0.216 ms) and lean bench
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02ac_*
timeout 120 build/igemm_trace 128 128 112 16 3 0 > $O/r02ac_trace_conv2_2_bf16_pool0.txt 2>&1; sed -n 1,2p $O/r02ac_trace_conv2_2_bf16_pool0.txt; tail -52 $O/r02ac_trace_conv2_2_bf16_pool0.txt | head -12
timeout 120 build/igemm_trace 64 64 224 16 3 1 > $O/r02ac_trace_conv1_2_bf16_pool1.txt 2>&1; sed -n 1,2p $O/r02ac_trace_conv1_2_bf16_pool1.txt
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_variants.py -m gpu -q -x -p no:cacheprovider -k "sgeconv or variants or pool" 2>&1 | tail -2
run() { echo "== $MODEL $*" >> $O/r02ac_lean.log; env "$@" timeout 300 python bench.py --lean --model $MODEL >> $O/r02ac_lean.log 2>&1; }
for MODEL in vgg16 resnet50 mobilenet_v1; do run A=1; done
grep -E "==|lean|rror|timed" $O/r02ac_lean.log
