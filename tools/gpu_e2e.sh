#!/bin/bash
# spmma tests of the tree + the driver's bench line with both e2e legs
T=${1:-e2e}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "spmma" --maxfail=10 --tb=short -p no:cacheprovider > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 $O/${T}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
tail -5 $O/${T}_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/e2e_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'])
print(json.dumps(d['e2e'], indent=1))
P
