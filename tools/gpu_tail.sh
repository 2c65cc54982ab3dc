#!/bin/bash
# tail-split trip: spmma tests, then the single-call sweep with and without the split (dev library switch)
T=${1:-tail}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_conv.py -m gpu -q -k "spmma or conv or full" --maxfail=10 --tb=short -p no:cacheprovider > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O/${T}_pytest.log
DEV=$PWD/sparsify.me_b200/lib_dev/libsparsifyme_b200.so
timeout 300 python tools/layer_sweep.py --tag split > $O/${T}_sweep_split.csv 2>&1
SPFY_LIB=$DEV SPFY_SPMMA_NO_TAIL_SPLIT=1 timeout 300 python tools/layer_sweep.py --tag nosplit > $O/${T}_sweep_nosplit.csv 2>&1
SPFY_LIB=$DEV timeout 300 python tools/layer_sweep.py --tag devsplit > $O/${T}_sweep_devsplit.csv 2>&1
paste -d' ' <(cut -d, -f2-4,6 $O/${T}_sweep_split.csv) <(cut -d, -f6 $O/${T}_sweep_nosplit.csv) <(cut -d, -f6 $O/${T}_sweep_devsplit.csv)
