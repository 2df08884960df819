#!/bin/bash
# round 2u: one bulk copy per k-block for the BF16 filter planes; TMA small-copy rates; TensorGEMM cluster multicast re-test
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02u_*
timeout 120 build/store_rate 2>&1 | grep -E "bulk copy|tma load" > $O/r02u_tma_rates.txt; cat $O/r02u_tma_rates.txt
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_variants.py -m gpu -q -x -p no:cacheprovider -k "sgeconv or variants" 2>&1 | tail -3
timeout 120 build/igemm_trace 64 64 224 16 3 0 > $O/r02u_trace_conv1_2_bf16_pool0.txt 2>&1; sed -n 1,2p $O/r02u_trace_conv1_2_bf16_pool0.txt; sed -n 30,40p $O/r02u_trace_conv1_2_bf16_pool0.txt; tail -9 $O/r02u_trace_conv1_2_bf16_pool0.txt
run() { echo "== $MODEL $*" >> $O/r02u_lean.log; env "$@" timeout 300 python bench.py --lean --model $MODEL >> $O/r02u_lean.log 2>&1; }
MODEL=vgg16; run A=1; run FCUDA_GEMM_CLUSTER=2; run FCUDA_IGEMM_TMA_OUT=0
MODEL=resnet50; run A=1; run FCUDA_IGEMM_TMA_OUT=0
MODEL=mobilenet_v1; run A=1
grep -E "==|lean|Error|error|timed" $O/r02u_lean.log
