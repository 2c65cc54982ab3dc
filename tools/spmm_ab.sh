#!/bin/bash
# A/B of the unstructured SpMM between gpurun_ab/lib_prev.so and the current build on a few shapes
for v in prev cur prev cur; do
  if [ $v = prev ]; then export SPFY_LIB=$PWD/gpurun_ab/lib_prev.so; else unset SPFY_LIB; fi
  for c in "64 576 12544 32 0.95" "64 576 12544 32 0.9" "256 2304 784 32 0.95" "512 4608 196 32 0.9" "128 1152 3136 32 0.5"; do
    echo -n "$v $c : "; python tools/spmm_one.py $c | tail -1
  done
done
