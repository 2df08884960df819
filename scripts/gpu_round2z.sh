#!/bin/bash
# round 2z: flattened pixel boxes for small images on the generic implicit-GEMM path
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02z_*
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > $O/r02z_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02z_pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" $O/r02z_pytest_gpu.log | tail -25
run() { echo "== $MODEL $*" >> $O/r02z_lean.log; env "$@" timeout 300 python bench.py --lean --model $MODEL >> $O/r02z_lean.log 2>&1; }
for MODEL in resnet50 mobilenet_v1 vgg16; do run A=1; done
grep -E "==|lean|Error|error|timed" $O/r02z_lean.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02z_resnet50_launches.csv python bench.py --model resnet50 --steps 1 --warmup 3 --no-graph --lean > $O/r02z_resnet50_launches.stdout 2>&1
python scripts/summarize_launches.py $O/r02z_resnet50_launches.csv $O/r02z_resnet50_traffic.json > $O/r02z_resnet50_launches_summary.txt 2>&1
sed -n '/one Forward/,$p' $O/r02z_resnet50_launches_summary.txt | head -20
