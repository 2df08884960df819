#!/bin/bash
# Two GPUs of one box: the C++ NetGroup path (NCCL broadcast bound at run time) and the torchrun bench line.
TAG=${1:-r02zz}
O=gpurun_out
mkdir -p $O; rm -f $O/${TAG}_n2_*
nvidia-smi -L | head -4
NCCL_DEBUG=WARN timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q -p no:cacheprovider -k "net_group or weight_arena" 2>&1 | tail -6 > $O/${TAG}_n2_pytest_netgroup.log; cat $O/${TAG}_n2_pytest_netgroup.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 2>$O/${TAG}_n2_bench.err | tail -1 > $O/${TAG}_n2_bench.json
python - <<PY
import json
d=json.load(open('$O/${TAG}_n2_bench.json'))
print('n_gpus', d['n_gpus'], 'vgg16', round(d['value'],1), 'img/s e2e', round(d['e2e']['value'],1), 'checksums equal', d['rank_checksums']['equal'])
for w in d.get('workloads', []): print(w.get('model'), round(w.get('value',0),1), 'e2e', round(w.get('e2e',{}).get('value',0),1), (w.get('rank_checksums') or {}).get('equal'), w.get('error'))
PY
tail -3 $O/${TAG}_n2_bench.err
