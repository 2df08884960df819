#!/bin/bash
# Runs on the GPU box (one GPU): the round-2 evidence set that is copied into profiles/ —
# tests, the default bench line (all workloads, all legs), launch lists with DRAM traffic, ncu --set full of the top kernels.
TAG=${1:-r02zz}
O=gpurun_out
mkdir -p $O; rm -f $O/${TAG}_*
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -5 > $O/${TAG}_pytest_gpu.log; cat $O/${TAG}_pytest_gpu.log
timeout 900 python bench.py 2>$O/${TAG}_bench_default.err | tail -1 > $O/${TAG}_bench_default.json
python - <<PY
import json
d=json.load(open('$O/${TAG}_bench_default.json'))
def show(m, w):
    r=w.get('roofline') or {}
    print(m, round(w['value'],1),'img/s dev', round(w['ms_per_step'],3),'ms | e2e', round(w['e2e']['value'],1), '| parity', (w.get('parity') or {}).get('parity_max_rel'),
          '| top:', (r.get('kernel') or '?').split()[0], round(r.get('achieved',0),1), r.get('unit'), 'frac', round(r.get('frac',0),3), 'share', round(r.get('share_of_step',0),3),
          '| cpu', (w.get('cpu_baseline') or {}).get('value'))
show('vgg16', d)
for w in d.get('workloads', []):
    if 'error' in w: print(w)
    else: show(w['model'], w)
print('clocks', d.get('clocks'))
PY
for M in vgg16 resnet50 mobilenet_v1; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/${TAG}_${M}_launches.csv \
      python bench.py --model $M --steps 1 --warmup 3 --no-graph --lean > $O/${TAG}_${M}_launches.stdout 2>&1
  python scripts/summarize_launches.py $O/${TAG}_${M}_launches.csv $O/${TAG}_${M}_traffic.json > $O/${TAG}_${M}_launches_summary.txt 2>&1
  echo "=== $M"; sed -n '/one Forward/,$p' $O/${TAG}_${M}_launches_summary.txt | head -18
done
python - <<PY
import json
out={}
for m in ('vgg16','resnet50','mobilenet_v1'):
    try: out[m]=json.load(open('$O/${TAG}_%s_traffic.json' % m)).get(m) or json.load(open('$O/${TAG}_%s_traffic.json' % m))
    except Exception as e: out[m]={'error': str(e)}
json.dump(out, open('$O/${TAG}_kernel_traffic.json','w'), indent=1)
PY
for K in conv_igemm_kernel tensor_gemm_ts_kernel wino_input_kernel wino_output_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 12 -c 4 -o $O/${TAG}_vgg16_$K -f \
      python bench.py --model vgg16 --steps 1 --warmup 3 --no-graph --lean > $O/${TAG}_vgg16_${K}_ncu.stdout 2>&1
  ncu -i $O/${TAG}_vgg16_$K.ncu-rep --page raw --csv 2>/dev/null | python scripts/summarize_ncu_raw.py > $O/${TAG}_vgg16_${K}_summary.txt 2>&1
  echo "=== $K"; grep -E "Kernel Name|gpu__time_duration|dram__bytes_(read|write).sum |sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|gpu__dram_throughput|sm__warps_active" $O/${TAG}_vgg16_${K}_summary.txt | cut -c1-160 | head -28
  rm -f $O/${TAG}_vgg16_$K.ncu-rep
done
timeout 600 ncu --set full --clock-control none -k regex:conv_igemm_kernel -s 60 -c 6 -o $O/${TAG}_resnet50_igemm -f \
    python bench.py --model resnet50 --steps 1 --warmup 3 --no-graph --lean > $O/${TAG}_resnet50_igemm_ncu.stdout 2>&1
ncu -i $O/${TAG}_resnet50_igemm.ncu-rep --page raw --csv 2>/dev/null | python scripts/summarize_ncu_raw.py > $O/${TAG}_resnet50_conv_igemm_kernel_summary.txt 2>&1
rm -f $O/${TAG}_resnet50_igemm.ncu-rep
grep -E "Kernel Name|gpu__time_duration|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|gpu__dram_throughput" $O/${TAG}_resnet50_conv_igemm_kernel_summary.txt | cut -c1-160 | head -24
ls -la $O/${TAG}_* | awk '{print $5, $9}'
