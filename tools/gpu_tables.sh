#!/bin/bash
# the secondary tables of a round: COO sweep (ResNet-34), the reference's compare.csv table (ResNet-50), dense GEMM probe
T=${1:-tab}; O=gpurun_out; mkdir -p $O
timeout 900 python tools/spmm_sweep.py --csv resnet34.csv --no-cusparse --tag $T > $O/${T}_spmm_sweep.csv 2> $O/${T}_spmm_sweep.err; echo "sweep rc=$?"
grep "^#" $O/${T}_spmm_sweep.csv
timeout 900 examples/bin/compare datasets/resnet50.csv > $O/${T}_compare.csv 2> $O/${T}_compare.err; echo "compare rc=$?"
tail -4 $O/${T}_compare.csv
timeout 300 python tools/gemm_probe.py --time --dtype f32 > $O/${T}_gemm_f32.txt 2>&1
timeout 300 python tools/gemm_probe.py --time --dtype f16 > $O/${T}_gemm_f16.txt 2>&1
