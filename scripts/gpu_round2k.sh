#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02k_*
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $O/r02k_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02k_pytest_gpu.log
grep -E "^FAILED|passed|failed|rc=" $O/r02k_pytest_gpu.log | tail -20
for m in vgg16 resnet50 mobilenet_v1; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file $O/r02k_${m}_launches.csv python bench.py --lean --no-graph --model $m --steps 2 --warmup 3 > $O/r02k_${m}_ncu.log 2>&1
done
ls -la $O/r02k_*
