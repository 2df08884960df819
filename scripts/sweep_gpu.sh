#!/bin/bash
# Quick characterisation sweeps on the GPU box (lean bench = device-resident leg only).
OUT=gpurun_out
for mb in 12 24 48 96 192 0; do
  echo "== l2-chunk-mb $mb"; python bench.py --lean --steps 10 --l2-chunk-mb $mb 2>/dev/null | tail -1
done
echo "== tf32 (1x) chunk 48"; python bench.py --lean --steps 10 --precision tf32 2>/dev/null | tail -1
echo "== no graph"; python bench.py --lean --steps 10 --no-graph 2>/dev/null | tail -1
echo "== no fusion"; python bench.py --lean --steps 10 --no-fusion 2>/dev/null | tail -1
for m in resnet50 mobilenet_v1; do echo "== $m"; python bench.py --lean --steps 10 --model $m 2>/dev/null | tail -1; done
