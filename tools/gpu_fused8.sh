#!/bin/bash
# bench --strong at N GPUs (fused gather beside GEMM + ncclAllGather)
T=${1:-fg8}; N=${2:-8}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
timeout 900 $TR bench.py --gpus $N --strong --steps 5 --warmup 3 > $O/${T}_strong.json 2> $O/${T}_strong.err
echo "strong rc=$?"
tail -3 $O/${T}_strong.err; cat $O/${T}_strong.json
