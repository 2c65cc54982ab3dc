#!/bin/bash
O=gpurun_out; T=${1:-wh}
export SPFY_LIB=$PWD/sparsify.me_b200/lib_dev/libsparsifyme_b200.so
for w in 1 0; do
  echo "== SPFY_GEMM_WRITE_HI=$w" >> $O/${T}_wh.log
  SPFY_GEMM_WRITE_HI=$w timeout 200 python tools/gemm_probe.py --dtype f32 2>&1 | grep "ta=1 tb=0 prec=0" >> $O/${T}_wh.log
  SPFY_GEMM_WRITE_HI=$w timeout 200 python tools/gemm_probe.py --dtype f32 --time 2>&1 | grep "prec=0" >> $O/${T}_wh.log
done
cat $O/${T}_wh.log
