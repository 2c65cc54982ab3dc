#!/bin/bash
# A/B of sliced residency (dev build: SPFY_SPMMA_NO_SLICES switches it off) + the spmma parity tests on the release build
T=${1:-sl}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_conv.py -m gpu -x -q -k "spmma or plan or fullsize or conv" > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${T}_pytest.log
export SPFY_LIB=$PWD/sparsify.me_b200/lib_dev/libsparsifyme_b200.so
for sw in 1 0; do
  if [ $sw = 1 ]; then export SPFY_SPMMA_NO_SLICES=1; else unset SPFY_SPMMA_NO_SLICES; fi
  echo "== NO_SLICES=$sw" >> $O/${T}_ab.log
  timeout 300 python tools/layer_sweep.py --only 1024,256,25088 --plan --tag ns$sw >> $O/${T}_ab.log 2>&1
done
cat $O/${T}_ab.log
