#!/bin/bash
# flexible GPU visit; steps chosen by words in $2..: test bench prune spmm sweep ref ncu_spmm ncu_spmma ncu_prune dbg
TAG=${1:-r1}; shift
OUT=gpurun_out
mkdir -p $OUT
for step in "$@"; do
case $step in
test) python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu_$TAG.log;;
bench) python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?";;
prune) python tools/prune_probe.py --tag $TAG > $OUT/prune_probe_$TAG.csv 2>&1; echo "prune probe rc=$?";;
spmm) python tools/spmm_sweep.py --tag $TAG > $OUT/spmm_sweep_$TAG.csv 2>&1; echo "spmm sweep rc=$?";;
sweep) python tools/layer_sweep.py --plan --tag $TAG > $OUT/sweep_$TAG.csv 2>&1; echo "sweep rc=$?";;
ref) python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?";;
dbg) for d in 2 8 10; do SPFY_SPMMA_DEBUG=$d python tools/layer_sweep.py --tag dbg$d > $OUT/sweep_${TAG}_dbg$d.csv 2>&1; echo "dbg$d rc=$?"; done;;
plandbg) for d in 32 160 288 416; do SPFY_SPMMA_DEBUG=$d python tools/layer_sweep.py --plan-only --tag dbg$d > $OUT/plan_${TAG}_dbg$d.txt 2>&1; echo "plan dbg$d rc=$?"; done;;
pfsweep) for d in 0 2 4 6 8 12 16 24; do SPFY_SPMMA_PF=$d python tools/layer_sweep.py --plan-only --tag pf$d > $OUT/plan_${TAG}_pf$d.txt 2>&1; echo "plan pf$d rc=$?"; grep "^#" $OUT/plan_${TAG}_pf$d.txt | sed 's/{[^}]*}//'; done;;
g1exp) SPFY_SPMMA_FORCE_G1=1 python tools/layer_sweep.py --plan-only --tag g1 > $OUT/plan_${TAG}_g1.txt 2>&1; grep "^#" $OUT/plan_${TAG}_g1.txt | sed 's/{[^}]*}//';;
stexp) for c in 1 2; do SPFY_SPMMA_STAGES=$c python tools/layer_sweep.py --plan-only --tag st$c > $OUT/plan_${TAG}_st$c.txt 2>&1; grep "^#" $OUT/plan_${TAG}_st$c.txt | sed 's/{[^}]*}//'; done;;
examples)
  ( cd examples && ./bin/sparsify 12544 147 && ./bin/spmma 64 25088 576 32 && ./bin/spmm 64 64 128 4 && ./bin/batched_coo 64 196 576 8 && ./bin/gemm 64 196 576 8 && ./bin/sweep ../datasets/resnet18.csv weights ) > $OUT/examples_$TAG.log 2>&1; echo "examples rc=$?"; tail -5 $OUT/examples_$TAG.log;;
mg)
  NG=${NGPUS:-2}
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 50 --warmup 5 > $OUT/bench_mg${NG}_$TAG.json 2> $OUT/bench_mg${NG}_$TAG.err; echo "mg rc=$?"; tail -c 600 $OUT/bench_mg${NG}_$TAG.json;;
tileprobe)
  bash tests/golden/run_tile_probe.sh > $OUT/tileprobe_$TAG.log 2>&1; echo "tile probe rc=$?"
  python tests/golden/make_tile_probe.py fixtures $OUT/tileprobe tests/golden > $OUT/tilefix_$TAG.log 2>&1; echo "tile fixtures rc=$?"; tail -15 $OUT/tilefix_$TAG.log
  mkdir -p $OUT/golden_tile && cp tests/golden/tile_*.npz $OUT/golden_tile/
  cp $OUT/tileprobe/special.in.bin $OUT/tileprobe/special.tile.bin $OUT/golden_tile/ 2>/dev/null
  rm -rf $OUT/tileprobe;;  # 60 MB of regenerable probe matrices: gpurun_out/ must stay under 64 MiB
compare) ( cd examples && ./bin/compare ../datasets/resnet50.csv > ../$OUT/compare_$TAG.csv 2> ../$OUT/compare_$TAG.err ); echo "compare rc=$?"; tail -2 $OUT/compare_$TAG.csv;;
thr) for a in "8192 8192 0.5" "8192 8192 0.1" "16384 8192 0.5 f16" "401408 147 0.1"; do timeout 100 python tools/thr_one.py $a | tail -1; done > $OUT/thr_one_$TAG.log 2>&1; cat $OUT/thr_one_$TAG.log; timeout 120 python tools/thr_probe.py 2>&1 | grep "C ABI" | tee -a $OUT/thr_one_$TAG.log;;
pcie) python tools/pcie_probe.py > $OUT/pcie_$TAG.csv 2>&1; echo "pcie rc=$?"; cat $OUT/pcie_$TAG.csv;;
ktable) python tools/kernel_table.py --tag $TAG > $OUT/kernels_$TAG.csv 2> $OUT/kernels_$TAG.err; echo "ktable rc=$?"; tail -3 $OUT/kernels_$TAG.err;;
ab) for r in 1 2 3; do for v in prev cur; do
      if [ $v = prev ]; then export SPFY_LIB=$PWD/gpurun_ab/lib_prev.so; else unset SPFY_LIB; fi
      python tools/layer_sweep.py --plan-only --tag $v$r 2>&1 | grep "^#" | sed 's/{[^}]*}//' | tr '\n' ' '; echo; done; done; unset SPFY_LIB;;
mgcoo)
  NG=${NGPUS:-2}
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --workload coo --gpus $NG --steps 5 --warmup 2 > $OUT/bench_coo_mg${NG}_$TAG.json 2> $OUT/bench_coo_mg${NG}_$TAG.err; echo "mgcoo rc=$?"; tail -c 400 $OUT/bench_coo_mg${NG}_$TAG.json;;
mg152)
  NG=${NGPUS:-2}
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29515 bench.py --csv resnet152.csv --batch 32 --gpus $NG --steps 30 --warmup 5 --no-cpu --no-prune-large > $OUT/bench_r152_mg${NG}_$TAG.json 2> $OUT/bench_r152_mg${NG}_$TAG.err; echo "mg152 rc=$?"; tail -c 300 $OUT/bench_r152_mg${NG}_$TAG.json;;
ncu_spmm)
  CMD="python tools/spmm_one.py 256 2304 784 32 0.9"
  $CMD > $OUT/spmm_one_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:spmm_csr -s 1 -c 1 -o $OUT/prof_spmm_$TAG -f $CMD > $OUT/ncu_spmm_$TAG.log 2>&1; echo "ncu spmm rc=$?";;
ncu_walk)
  CMD="python tools/spmm_one.py 256 2304 784 32 0.5"
  $CMD > $OUT/walk_one_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:spmm_dense_walk -s 1 -c 1 -o $OUT/prof_walk_$TAG -f $CMD > $OUT/ncu_walk_$TAG.log 2>&1; echo "ncu walk rc=$?"; cat $OUT/walk_one_$TAG.log;;
ncu_thr)
  CMD="python tools/thr_one.py 8192 8192 0.5"
  ncu --set full --clock-control none --import-source on -k regex:threshold_compact -s 1 -c 1 -o $OUT/prof_thr_$TAG -f $CMD > $OUT/ncu_thr_$TAG.log 2>&1; echo "ncu thr rc=$?";;
ncu_tile)
  CMD="python tools/kernel_table.py --reps 1"
  ncu --set full --clock-control none --import-source on -k regex:prune24_tile -s 1 -c 1 -o $OUT/prof_tile_$TAG -f $CMD > $OUT/ncu_tile_$TAG.log 2>&1; echo "ncu tile rc=$?";;
ncu_spmma)
  CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-prune-large"
  $CMD > $OUT/plain_$TAG.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmma_kernel|prune24' -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_l_$TAG.log 2>&1; echo "launch list rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:spmma_kernel -s 18 -c 6 -o $OUT/prof_spmma_$TAG -f $CMD > $OUT/ncu_s_$TAG.log 2>&1; echo "ncu spmma rc=$?";;
ncu_prune)
  CMD="python tools/prune_probe.py --reps 1 --only-large"
  $CMD > $OUT/plain_prune_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:prune24_fast -s 3 -c 1 -o $OUT/prof_prune_$TAG -f $CMD > $OUT/ncu_p_$TAG.log 2>&1; echo "ncu prune rc=$?";;
esac
done
