#!/bin/bash
T=${1:-sc}; O=gpurun_out; mkdir -p $O
timeout 600 oracle/_ref/cusparselt_ref sweep datasets/resnet50.csv 32 > $O/${T}_cusparselt_layers.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-prune-large --per-layer > $O/${T}_bench.json 2> $O/${T}_ours_layers.txt
grep -c layer $O/${T}_cusparselt_layers.txt $O/${T}_ours_layers.txt
