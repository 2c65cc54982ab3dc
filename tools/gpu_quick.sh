#!/bin/bash
# short trip: tensor-route tests + GEMM timing probe (+ optional env-switch runs with the dev library: "VAR=val ...")
T=${1:-q}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_headers.py -q --maxfail=20 --tb=short -p no:cacheprovider > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?" > $O/${T}_status.txt
timeout 300 python tools/gemm_probe.py --dtype f32 --time > $O/${T}_probe_time_f32.log 2>&1
shift
for kv in "$@"; do
  echo "== $kv" >> $O/${T}_dbg.log
  env SPFY_LIB=$PWD/sparsify.me_b200/lib_dev/libsparsifyme_b200.so $kv timeout 200 python tools/gemm_probe.py --dtype f32 --time 2>&1 | grep "prec=0" >> $O/${T}_dbg.log
done
timeout 600 python tools/spmm_sweep.py --csv resnet34.csv --reps 3 --no-cusparse --tag $T > $O/${T}_spmm_sweep.csv 2> $O/${T}_spmm_sweep.err
cat $O/${T}_status.txt; tail -3 $O/${T}_pytest.log; grep "prec=0" $O/${T}_probe_time_f32.log; cat $O/${T}_dbg.log 2>/dev/null; tail -3 $O/${T}_spmm_sweep.csv
