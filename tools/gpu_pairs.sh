#!/bin/bash
# CTA-pair dense GEMM: parity tests, then timing with and without pairs (dev build switch)
T=${1:-pr}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tensor.py -m gpu -x -q -p no:cacheprovider > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/${T}_pytest.log
export SPFY_LIB=$PWD/sparsify.me_b200/lib_dev/libsparsifyme_b200.so
for sw in 1 0; do
  if [ $sw = 1 ]; then export SPFY_GEMM_NO_PAIRS=1; else unset SPFY_GEMM_NO_PAIRS; fi
  echo "== NO_PAIRS=$sw" >> $O/${T}_time.log
  timeout 300 python tools/gemm_probe.py --time --dtype f16 >> $O/${T}_time.log 2>&1
  timeout 300 python tools/gemm_probe.py --dtype f16 >> $O/${T}_probe_$sw.log 2>&1
done
cat $O/${T}_time.log
