#!/bin/bash
# what the driver runs for a scaling point: reference arm and our arm under torchrun at N GPUs, plus the multi-GPU tests
T=${1:-sc}; N=${2:-2}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555"
timeout 600 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${T}_pytest.log
timeout 900 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/${T}_ref_n$N.json 2> $O/${T}_ref_n$N.err; echo "ref rc=$?"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/${T}_bench_n$N.json 2> $O/${T}_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
for f in ['$O/${T}_ref_n$N.json','$O/${T}_bench_n$N.json']:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('/')[-1], d.get('impl'), d.get('value'), d.get('n_gpus'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('cores'))
PY
