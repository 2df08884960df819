#!/bin/bash
# Runs on the GPU box: the evidence set that is copied into profiles/ (tests, bench lines, launch lists, ncu captures).
TAG=${1:-r01z}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/${TAG}_pytest_gpu.log
timeout 120 build/gemm_selftest --bench > $O/${TAG}_gemm_selftest.log 2>&1
for M in vgg16 resnet50 mobilenet_v1; do
  timeout 600 python bench.py --model $M 2>$O/${TAG}_bench_$M.err | tail -1 > $O/${TAG}_bench_$M.json
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>$O/${TAG}_bench_reference.err | tail -1 > $O/${TAG}_bench_reference_vgg16.json
scripts/launches_gpu.sh $TAG vgg16 resnet50 mobilenet_v1 > $O/${TAG}_launches.log 2>&1
scripts/profile_gpu.sh vgg16 $TAG conv_igemm_kernel tensor_gemm_ts_kernel wino_input_kernel wino_output_kernel > $O/${TAG}_ncu.log 2>&1
rm -f $O/${TAG}_*.ncu-rep.tmp
cat $O/${TAG}_pytest_gpu.log; tail -1 $O/${TAG}_gemm_selftest.log
for M in vgg16 resnet50 mobilenet_v1; do python -c "
import json,sys
d=json.load(open('$O/${TAG}_bench_$M.json'))
r=d['roofline']
print('$M', round(d['value'],1),'img/s dev', round(d['e2e']['value'],1),'e2e', d['cpu_baseline']['value'], 'cpu', '| top:',r['kernel'].split()[0], round(r['achieved'],1), r['unit'], round(r['frac'],3), 'share', round(r['share_of_step'],3))
"; done
cut -c1-300 $O/${TAG}_bench_reference_vgg16.json
