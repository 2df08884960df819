#!/bin/bash
# Runs on the GPU box: per-launch device times (ncu, cold-cache, serialised) of one bench step for each model given.
# Usage: scripts/launches_gpu.sh <tag> <model> [<model> ...]      outputs gpurun_out/<tag>_<model>_launches_summary.txt
TAG=$1; shift
mkdir -p gpurun_out
for M in "$@"; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${TAG}_${M}_launches.csv \
      python bench.py --model $M --steps 1 --warmup 3 --no-graph --lean > gpurun_out/${TAG}_${M}_launches.stdout 2>&1
  python scripts/summarize_launches.py gpurun_out/${TAG}_${M}_launches.csv gpurun_out/${TAG}_${M}_traffic.json > gpurun_out/${TAG}_${M}_launches_summary.txt 2>&1
  echo "=== $M"; sed -n '/one Forward/,$p' gpurun_out/${TAG}_${M}_launches_summary.txt | head -16
done
