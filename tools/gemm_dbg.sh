#!/bin/bash
# timing experiments on the 3xTF32 kernel with pieces switched off (dev build; results are garbage by design)
O=gpurun_out; T=${1:-dbg}
export SPFY_LIB=$PWD/sparsify.me_b200/lib_dev/libsparsifyme_b200.so
for d in 0 1 2 4 8 6 14 15; do
  echo "== SPFY_GEMM_DEBUG=$d" >> $O/${T}_gemm_dbg.log
  SPFY_GEMM_DEBUG=$d timeout 200 python tools/gemm_probe.py --dtype f32 --time 2>&1 | grep "prec=0" >> $O/${T}_gemm_dbg.log
done
for st in 2 4; do
  echo "== SPFY_GEMM_STAGES=$st" >> $O/${T}_gemm_dbg.log
  SPFY_GEMM_STAGES=$st timeout 200 python tools/gemm_probe.py --dtype f32 --time 2>&1 | grep "prec=0" >> $O/${T}_gemm_dbg.log
done
cat $O/${T}_gemm_dbg.log
