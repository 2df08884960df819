#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02v_*
timeout 120 build/store_rate 2>&1 | grep -E "bulk copy" > $O/r02v_tma_rates.txt; cat $O/r02v_tma_rates.txt
timeout 120 build/igemm_trace 64 64 224 16 3 0 > $O/r02v_trace_conv1_2_bf16_pool0.txt 2>&1; sed -n 1,2p $O/r02v_trace_conv1_2_bf16_pool0.txt; sed -n 30,36p $O/r02v_trace_conv1_2_bf16_pool0.txt; tail -30 $O/r02v_trace_conv1_2_bf16_pool0.txt
