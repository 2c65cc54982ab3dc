#!/bin/bash
# the driver's scaling point at N GPUs on the final tree: weak headline (+ reference arm on rank 0), COO workload
T=${1:-s8}; N=${2:-8}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29566"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
timeout 900 $TR bench.py --gpus $N --workload coo --steps 5 --warmup 2 > $O/${T}_coo.json 2> $O/${T}_coo.err; echo "coo rc=$?"
python - <<PY
import json
for f in ['$O/${T}_bench.json','$O/${T}_coo.json']:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('/')[-1], d.get('value'), d.get('n_gpus'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'))
PY
