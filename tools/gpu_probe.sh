#!/bin/bash
# measurement visit: parity tests, bench (both arms), per-shape spmma sweep, unstructured sweep, prune probe
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu_$TAG.log
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python tools/prune_probe.py --tag $TAG > $OUT/prune_probe_$TAG.csv 2>&1; echo "prune probe rc=$?"
python tools/spmm_sweep.py --tag $TAG > $OUT/spmm_sweep_$TAG.csv 2>&1; echo "spmm sweep rc=$?"
if [ "$2" = "full" ]; then
python tools/layer_sweep.py --plan --tag $TAG > $OUT/sweep_$TAG.csv 2>&1; echo "sweep rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"
fi
