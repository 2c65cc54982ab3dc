#!/bin/bash
# fused output gather: single-GPU replica test, the 2-GPU worker test, bench --strong at N GPUs
T=${1:-fg}; N=${2:-2}; O=gpurun_out; mkdir -p $O
nvidia-smi topo -m > $O/${T}_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "replicated or plan" -p no:cacheprovider > $O/${T}_pytest1.log 2>&1
echo "pytest replicas rc=$?" > $O/${T}_status.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -q --tb=short -p no:cacheprovider > $O/${T}_pytest.log 2>&1
echo "pytest multi rc=$?" >> $O/${T}_status.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
timeout 900 $TR bench.py --gpus $N --strong --steps 5 --warmup 3 > $O/${T}_strong.json 2> $O/${T}_strong.err
echo "strong rc=$?" >> $O/${T}_status.txt
cat $O/${T}_status.txt; tail -5 $O/${T}_pytest1.log; tail -15 $O/${T}_pytest.log; tail -5 $O/${T}_strong.err; cat $O/${T}_strong.json
