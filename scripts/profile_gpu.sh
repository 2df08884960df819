#!/bin/bash
# Runs on the GPU box (under gpurun): ncu launch list of one bench step and a full capture of the dominant kernel.
# Usage: scripts/profile_gpu.sh <model> <tag>      outputs under gpurun_out/
set -u
MODEL=${1:-vgg16}
TAG=${2:-r01}
OUT=gpurun_out
mkdir -p $OUT
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_${MODEL}_launches.csv \
    python bench.py --model $MODEL --steps 1 --warmup 3 --no-graph --lean > $OUT/${TAG}_${MODEL}_launches.stdout 2>&1
python scripts/summarize_launches.py $OUT/${TAG}_${MODEL}_launches.csv > $OUT/${TAG}_${MODEL}_launches_summary.txt 2>&1
# the dominant kernel, full set, 3 launches from the middle of a warmed-up step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tensor_gemm_kernel -s 60 -c 3 \
    -o $OUT/${TAG}_${MODEL}_tensor_gemm python bench.py --model $MODEL --steps 1 --warmup 3 --no-graph --lean \
    > $OUT/${TAG}_${MODEL}_ncu_full.stdout 2>&1
ncu -i $OUT/${TAG}_${MODEL}_tensor_gemm.ncu-rep --page raw --csv 2>/dev/null | python scripts/summarize_ncu_raw.py \
    > $OUT/${TAG}_${MODEL}_tensor_gemm_summary.txt 2>&1
tail -30 $OUT/${TAG}_${MODEL}_launches_summary.txt
cat $OUT/${TAG}_${MODEL}_tensor_gemm_summary.txt
