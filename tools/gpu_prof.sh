#!/bin/bash
# profiling trip (round 2): launch list of one bench step, ncu --set full of the plan's spmma launches, of the 3xTF32
# GEMM behind the COO SpMM and of the blocked-ELL expand + GEMM.  Every ncu command runs plain first.
T=${1:-r02}; O=gpurun_out; mkdir -p $O
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-prune-large"
$CMD > $O/${T}_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv $CMD > $O/${T}_ncu_l.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:spmma_kernel -s 18 -c 6 -o $O/${T}_prof_spmma -f $CMD > $O/${T}_ncu_s.log 2>&1
echo "ncu spmma rc=$?"
CMD2="python tools/spmm_one.py 256 2304 784 32 0.9"
$CMD2 > $O/${T}_plain_spmm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tcgemm -s 2 -c 1 -o $O/${T}_prof_tcgemm -f $CMD2 > $O/${T}_ncu_g.log 2>&1
echo "ncu tcgemm rc=$?"
CMD3="python tools/spmm_one.py 64 576 12544 32 0.9"
$CMD3 > $O/${T}_plain_spmm64.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tcgemm -s 2 -c 1 -o $O/${T}_prof_tcgemm64 -f $CMD3 > $O/${T}_ncu_g64.log 2>&1
echo "ncu tcgemm64 rc=$?"
CMD4="examples/bin/spmm 3136 128 1152 32"
$CMD4 > $O/${T}_plain_bell.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bell_expand -s 1 -c 1 -o $O/${T}_prof_bell -f $CMD4 > $O/${T}_ncu_b.log 2>&1
echo "ncu bell rc=$?"
ls -la $O/${T}_prof_*.ncu-rep
