#!/bin/bash
# Runs on the GPU box (under gpurun): ncu launch list of one bench step and full captures of the dominant kernels.
# Usage: scripts/profile_gpu.sh <model> <tag> [kernel-regex ...]      outputs under gpurun_out/
set -u
MODEL=${1:-vgg16}
TAG=${2:-r01}
shift 2
KERNELS=${@:-tensor_gemm_kernel}
OUT=gpurun_out
mkdir -p $OUT
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_${MODEL}_launches.csv \
    python bench.py --model $MODEL --steps 1 --warmup 3 --no-graph --lean > $OUT/${TAG}_${MODEL}_launches.stdout 2>&1
python scripts/summarize_launches.py $OUT/${TAG}_${MODEL}_launches.csv > $OUT/${TAG}_${MODEL}_launches_summary.txt 2>&1
tail -25 $OUT/${TAG}_${MODEL}_launches_summary.txt
for K in $KERNELS; do
  # full set, 2 launches after the warm-up forwards
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 12 -c 2 \
      -o $OUT/${TAG}_${MODEL}_$K -f python bench.py --model $MODEL --steps 1 --warmup 3 --no-graph --lean \
      > $OUT/${TAG}_${MODEL}_${K}_ncu.stdout 2>&1
  ncu -i $OUT/${TAG}_${MODEL}_$K.ncu-rep --page raw --csv 2>/dev/null | python scripts/summarize_ncu_raw.py \
      > $OUT/${TAG}_${MODEL}_${K}_summary.txt 2>&1
  echo "=== $K"; grep -E "^----|Kernel Name|gpu__time_duration|dram__bytes|lts__t_sector_hit|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|launch__grid|sm__warps_active|dram_throughput" $OUT/${TAG}_${MODEL}_${K}_summary.txt | cut -c1-170
done
