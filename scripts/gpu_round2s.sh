#!/bin/bash
# round 2s: explicit ld.shared / st.shared in the hot kernels, mbarrier suspend hint A/B (100 us / 1 us / poll)
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02s_*
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > $O/r02s_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02s_pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" $O/r02s_pytest_gpu.log | tail -25
run() { echo "== $MODEL $*" >> $O/r02s_lean.log; env "$@" timeout 300 python bench.py --lean --model $MODEL >> $O/r02s_lean.log 2>&1; }
for MODEL in vgg16 resnet50 mobilenet_v1; do run FCUDA_MBAR_SUSPEND_NS=100000; run FCUDA_MBAR_SUSPEND_NS=1000; run FCUDA_MBAR_SUSPEND_NS=0; done
grep -E "==|lean|Error|error|timed" $O/r02s_lean.log
timeout 120 build/igemm_trace 64 64 224 16 3 0 > $O/r02s_trace_conv1_2_bf16_pool0.txt 2>&1; sed -n 1,2p $O/r02s_trace_conv1_2_bf16_pool0.txt; sed -n 30,44p $O/r02s_trace_conv1_2_bf16_pool0.txt; tail -9 $O/r02s_trace_conv1_2_bf16_pool0.txt
