#!/bin/bash
# round 2t: pre-tiled BF16 filter planes fetched by 1-D bulk copies (un-swizzled core-matrix layout)
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02t_*
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -p no:cacheprovider -k "sgeconv and fp32split" > $O/r02t_quick.log 2>&1; tail -3 $O/r02t_quick.log
if ! grep -q " passed" $O/r02t_quick.log || grep -q "failed" $O/r02t_quick.log; then
  echo "== trying the swapped descriptor offsets"
  FCUDA_DEBUG_DESC_SWAP=1 timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -p no:cacheprovider -k "sgeconv and fp32split" 2>&1 | tail -3
fi
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > $O/r02t_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02t_pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" $O/r02t_pytest_gpu.log | tail -25
run() { echo "== $MODEL $*" >> $O/r02t_lean.log; env "$@" timeout 300 python bench.py --lean --model $MODEL >> $O/r02t_lean.log 2>&1; }
for MODEL in vgg16 resnet50 mobilenet_v1; do run A=1; done
grep -E "==|lean|Error|error|timed" $O/r02t_lean.log
timeout 120 build/igemm_trace 64 64 224 16 3 0 > $O/r02t_trace_conv1_2_bf16_pool0.txt 2>&1; sed -n 1,2p $O/r02t_trace_conv1_2_bf16_pool0.txt; sed -n 30,40p $O/r02t_trace_conv1_2_bf16_pool0.txt; tail -9 $O/r02t_trace_conv1_2_bf16_pool0.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02t_vgg16_launches.csv python bench.py --model vgg16 --steps 1 --warmup 3 --no-graph --lean > $O/r02t_vgg16_launches.stdout 2>&1
python scripts/summarize_launches.py $O/r02t_vgg16_launches.csv $O/r02t_vgg16_traffic.json > $O/r02t_vgg16_launches_summary.txt 2>&1
sed -n '/one Forward/,$p' $O/r02t_vgg16_launches_summary.txt | head -14
