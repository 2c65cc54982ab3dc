#!/bin/bash
# TILE prune trip: bit-exactness tests, then the large-matrix probe (two passes vs the fused kernel forced on)
T=${1:-tile}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_headers.py -m gpu -q -k "tile or prune or header or full or spmma_reference" --maxfail=10 --tb=short -p no:cacheprovider > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O/${T}_pytest.log
timeout 300 python tools/prune_probe.py --tile --tag release > $O/${T}_probe.csv 2>&1
DEV=$PWD/sparsify.me_b200/lib_dev/libsparsifyme_b200.so
SPFY_LIB=$DEV SPFY_TILE_FUSED_MAX=1000000000000 timeout 300 python tools/prune_probe.py --tile --tag fused-forced >> $O/${T}_probe.csv 2>&1
SPFY_LIB=$DEV SPFY_TILE_FUSED_MAX=0 timeout 300 python tools/prune_probe.py --tile --tag two-pass >> $O/${T}_probe.csv 2>&1
cat $O/${T}_probe.csv
