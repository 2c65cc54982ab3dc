#!/bin/bash
# multi-GPU trip: the NCCL gather test, then bench --strong (config 5) and the weak headline at this GPU count
T=${1:-mg}; N=${2:-2}; O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/${T}_gpus.txt 2>&1
nvidia-smi topo -m >> $O/${T}_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -q --tb=short -p no:cacheprovider > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?" > $O/${T}_status.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
timeout 900 $TR bench.py --gpus $N --strong --steps 5 --warmup 3 > $O/${T}_strong.json 2> $O/${T}_strong.err
echo "strong rc=$?" >> $O/${T}_status.txt
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --e2e-steps 2 > $O/${T}_weak.json 2> $O/${T}_weak.err
echo "weak rc=$?" >> $O/${T}_status.txt
timeout 900 $TR bench.py --gpus $N --workload coo --steps 5 --warmup 2 > $O/${T}_coo.json 2> $O/${T}_coo.err
echo "coo rc=$?" >> $O/${T}_status.txt
cat $O/${T}_status.txt; tail -3 $O/${T}_pytest.log; tail -3 $O/${T}_strong.err; head -c 600 $O/${T}_strong.json
