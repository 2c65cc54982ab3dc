#!/bin/bash
T=${1:-g2}; O=gpurun_out; mkdir -p $O
export SPFY_LIB=$PWD/sparsify.me_b200/lib_dev/libsparsifyme_b200.so
for w in 2.0 1.0 0.5 0.0; do
  echo "== SPFY_SPMMA_G2_MIN_WAVES=$w" >> $O/${T}_g2.log
  SPFY_SPMMA_G2_MIN_WAVES=$w timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-prune-large --per-layer 2>> $O/${T}_g2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('single_call', d['single_call']['ms_sum_over_layers'], 'plan', d['single_call']['plan_ms'])" >> $O/${T}_g2.log
done
cat $O/${T}_g2.log
