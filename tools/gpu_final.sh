#!/bin/bash
# final evidence of a round: every GPU test, smoke, the driver's bench line, the reference arm, launch list and ncu of the plan
T=${1:-fin}; O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -p no:cacheprovider > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --impl reference > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-prune-large"
$CMD > $O/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv $CMD > $O/${T}_ncu_l.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:spmma_kernel -s 18 -c 6 -o $O/${T}_prof_spmma -f $CMD > $O/${T}_ncu_s.log 2>&1
echo "ncu spmma rc=$?"
python tools/layer_sweep.py --plan --tag $T > $O/${T}_layer_sweep.csv 2>&1; echo "sweep rc=$?"
