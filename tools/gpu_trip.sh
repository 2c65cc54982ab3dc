#!/bin/bash
# one gpurun trip: probes first (so that a fault in a new kernel is localised), then the suites and the sweeps
TAG=${1:-r2a}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/${TAG}_gpu.txt 2>&1
for dt in f32 f16 bf16; do
  timeout 300 python tools/gemm_probe.py --dtype $dt > $O/${TAG}_probe_$dt.log 2>&1
  echo "probe $dt rc=$?" >> $O/${TAG}_status.txt
done
timeout 300 python tools/gemm_probe.py --dtype f32 --time > $O/${TAG}_probe_time_f32.log 2>&1
timeout 300 python tools/gemm_probe.py --dtype f16 --time > $O/${TAG}_probe_time_f16.log 2>&1
timeout 1500 python -m pytest tests -q -m gpu --maxfail=60 --tb=short -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_status.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "bench rc=$?" >> $O/${TAG}_status.txt
timeout 600 python tools/spmm_sweep.py --csv resnet34.csv --reps 3 --no-cusparse --tag $TAG > $O/${TAG}_spmm_sweep.csv 2> $O/${TAG}_spmm_sweep.err
echo "sweep rc=$?" >> $O/${TAG}_status.txt
timeout 600 examples/bin/compare datasets/resnet50.csv > $O/${TAG}_compare.csv 2> $O/${TAG}_compare.err
echo "compare rc=$?" >> $O/${TAG}_status.txt
timeout 600 oracle/_ref/cusparse_ref timebell datasets/resnet50.csv > $O/${TAG}_cusparse_bell.csv 2> $O/${TAG}_cusparse_bell.err
echo "timebell rc=$?" >> $O/${TAG}_status.txt
cat $O/${TAG}_status.txt
tail -5 $O/${TAG}_pytest.log
