#!/bin/bash
# round 2q: diagnostics — store / TMA-load rates per SM, igemm trace of conv1_2 (BF16x3, pooled / not), wino block order A/B
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02q_*
timeout 120 build/store_rate > $O/r02q_store_rate.txt 2>&1; cat $O/r02q_store_rate.txt
for pool in 0 1; do timeout 120 build/igemm_trace 64 64 224 16 3 $pool > $O/r02q_trace_conv1_2_bf16_pool$pool.txt 2>&1; sed -n 1,2p $O/r02q_trace_conv1_2_bf16_pool$pool.txt; sed -n 30,50p $O/r02q_trace_conv1_2_bf16_pool$pool.txt; tail -9 $O/r02q_trace_conv1_2_bf16_pool$pool.txt; done
timeout 120 build/igemm_trace 64 64 224 16 2 0 > $O/r02q_trace_conv1_2_tf32x3_pool0.txt 2>&1; sed -n 1,2p $O/r02q_trace_conv1_2_tf32x3_pool0.txt; sed -n 30,40p $O/r02q_trace_conv1_2_tf32x3_pool0.txt
run() { echo "== $MODEL $*" >> $O/r02q_lean.log; env "$@" timeout 300 python bench.py --lean --model $MODEL >> $O/r02q_lean.log 2>&1; }
MODEL=vgg16; run FCUDA_WINO_MLP=1; run FCUDA_WINO_MLP=2
grep -E "==|lean|Error|error|timed" $O/r02q_lean.log
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_variants.py -m gpu -q -x -p no:cacheprovider -k "wino or tensor_gemm or matches_oracle" 2>&1 | tail -3
