#!/bin/bash
# one gpurun trip: the whole GPU suite, the bench (both arms), the unstructured sweeps and the same-box comparators
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 1800 python -m pytest tests -q -m gpu --maxfail=40 --tb=short -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" > $O/${TAG}_status.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "bench rc=$?" >> $O/${TAG}_status.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
echo "reference rc=$?" >> $O/${TAG}_status.txt
timeout 600 python tools/spmm_sweep.py --csv resnet34.csv --reps 3 --no-cusparse --tag $TAG > $O/${TAG}_spmm_sweep.csv 2> $O/${TAG}_spmm_sweep.err
echo "sweep rc=$?" >> $O/${TAG}_status.txt
timeout 600 examples/bin/compare datasets/resnet50.csv > $O/${TAG}_compare.csv 2> $O/${TAG}_compare.err
echo "compare rc=$?" >> $O/${TAG}_status.txt
timeout 900 oracle/_ref/cusparse_ref timebell datasets/resnet50.csv > $O/${TAG}_cusparse_bell.csv 2> $O/${TAG}_cusparse_bell.err
echo "timebell rc=$?" >> $O/${TAG}_status.txt
cat $O/${TAG}_status.txt
tail -5 $O/${TAG}_pytest.log
tail -2 $O/${TAG}_spmm_sweep.csv
tail -1 $O/${TAG}_cusparse_bell.csv
