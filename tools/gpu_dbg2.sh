#!/bin/bash
T=${1:-dbg2}; O=gpurun_out; mkdir -p $O
export SPFY_LIB=$PWD/sparsify.me_b200/lib_dev/libsparsifyme_b200.so
for shapes in "64,64,401408" "256,64,401408" "512,128,100352"; do
  for d in 0 32 2 36 38; do
    echo "== shapes=$shapes DEBUG=$d" >> $O/${T}.log
    SPFY_SPMMA_DEBUG=$d timeout 300 python tools/layer_sweep.py --plan-only --plan --plan-shapes "$shapes" --tag d$d 2>&1 | grep "PLAN\|Error\|error" >> $O/${T}.log
  done
done
cat $O/${T}.log
