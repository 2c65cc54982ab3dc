#!/bin/bash
# One GPU-box visit: parity tests, bench, launch list, ncu captures of the two hot kernels.
# usage: tools/gpu_round.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -3 $OUT/pytest_gpu_$TAG.log
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python tools/layer_sweep.py --plan --tag $TAG > $OUT/sweep_$TAG.csv 2>&1; echo "sweep rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_l_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:spmma_kernel -s 9 -c 3 -o $OUT/prof_spmma_$TAG -f $CMD > $OUT/ncu_s_$TAG.log 2>&1
echo "ncu spmma rc=$?"
ncu --set full --clock-control none --import-source on -k regex:prune24 -s 3 -c 1 -o $OUT/prof_prune_$TAG -f $CMD > $OUT/ncu_p_$TAG.log 2>&1
echo "ncu prune rc=$?"
