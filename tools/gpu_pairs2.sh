#!/bin/bash
T=${1:-pr2}; O=gpurun_out; mkdir -p $O
export SPFY_LIB=$PWD/sparsify.me_b200/lib_dev/libsparsifyme_b200.so
for st in 2 3 4 7; do SPFY_GEMM_STAGES=$st timeout 120 python tools/gemm_one.py 784 256 2304 32 >> $O/${T}_stages.log 2>&1; done
for st in 2 3 5; do SPFY_GEMM_NO_PAIRS=1 SPFY_GEMM_STAGES=$st timeout 120 python tools/gemm_one.py 784 256 2304 32 >> $O/${T}_stages.log 2>&1; done
cat $O/${T}_stages.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tcgemm2_kernel -s 2 -c 1 -o $O/${T}_pair -f python tools/gemm_one.py 784 256 2304 32 > $O/${T}_ncu_pair.log 2>&1; echo "ncu pair rc=$?"
SPFY_GEMM_NO_PAIRS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tcgemm_kernel -s 2 -c 1 -o $O/${T}_single -f python tools/gemm_one.py 784 256 2304 32 > $O/${T}_ncu_single.log 2>&1; echo "ncu single rc=$?"
