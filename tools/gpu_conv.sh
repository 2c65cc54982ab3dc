#!/bin/bash
T=${1:-cv}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_conv.py -q --maxfail=30 --tb=short -p no:cacheprovider > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O/${T}_pytest.log
timeout 600 python tools/conv_sweep.py --tag $T > $O/${T}_conv_sweep.csv 2> $O/${T}_conv_sweep.err
cat $O/${T}_conv_sweep.csv; tail -3 $O/${T}_conv_sweep.err
