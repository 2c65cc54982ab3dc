#!/usr/bin/env python3
"""Hand-over timestamps of CTA 0 in ONE spfy_spmma launch (dev build: SPFY_LIB=.../lib_dev/..., SPFY_SPMMA_TRACE).
    SPFY_LIB=$PWD/sparsify.me_b200/lib_dev/libsparsifyme_b200.so python tools/trace_one.py M K N [out.csv]
Columns per role -- producer: t0 before / t1 after the wait for the unit's first free stage, t2 all loads of the unit issued;
MMA: t0 unit start, t1 accumulator slots free, t2 first stage landed, t3 MMAs + commits issued; epilogue warp 2 (per job):
t0 start, t1 accumulator ready, t2 drained + slot handed back, t3 staging buffer free, t4 store issued."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import torch  # noqa: E402

M, K, N = (int(v) for v in sys.argv[1:4])
out = sys.argv[4] if len(sys.argv) > 4 else "gpurun_out/trace.csv"
spfy = ge.load_package()
dev = torch.device("cuda:0")
w = (torch.rand(M, K, device=dev) * 2 - 1).half()
b = (torch.rand(K, N, device=dev) * 2 - 1).half()
d = torch.empty(M, N, dtype=torch.float16, device=dev)
comp = spfy.prune24(w)
for _ in range(3):
    spfy.spmma_compressed(comp, b, out=d)
torch.cuda.synchronize()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
flush.zero_()
torch.cuda.synchronize()
os.environ["SPFY_SPMMA_TRACE"] = out
spfy.spmma_compressed(comp, b, out=d)
torch.cuda.synchronize()
del os.environ["SPFY_SPMMA_TRACE"]
print(open(out).read())
