#!/bin/bash
# round 2y: pooled epilogue through swizzled shared memory
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02y_*
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > $O/r02y_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02y_pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" $O/r02y_pytest_gpu.log | tail -25
run() { echo "== $MODEL $*" >> $O/r02y_lean.log; env "$@" timeout 300 python bench.py --lean --model $MODEL >> $O/r02y_lean.log 2>&1; }
for MODEL in vgg16 resnet50 mobilenet_v1; do run A=1; done
grep -E "==|lean|Error|error|timed" $O/r02y_lean.log
timeout 120 build/igemm_trace 64 64 224 16 3 1 > $O/r02y_trace_conv1_2_bf16_pool1.txt 2>&1; sed -n 1,2p $O/r02y_trace_conv1_2_bf16_pool1.txt; tail -9 $O/r02y_trace_conv1_2_bf16_pool1.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02y_vgg16_launches.csv python bench.py --model vgg16 --steps 1 --warmup 3 --no-graph --lean > $O/r02y_vgg16_launches.stdout 2>&1
python scripts/summarize_launches.py $O/r02y_vgg16_launches.csv $O/r02y_vgg16_traffic.json > $O/r02y_vgg16_launches_summary.txt 2>&1
sed -n '/one Forward/,$p' $O/r02y_vgg16_launches_summary.txt | head -14
