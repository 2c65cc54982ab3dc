#!/usr/bin/env python3
"""One batched COO SpMM case, a few launches (the command ncu wraps).
    python tools/spmm_one.py M K n nb sparsity"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import torch  # noqa: E402

M, K, n, nb = (int(x) for x in sys.argv[1:5])
s = float(sys.argv[5])
spfy = ge.load_package()
dev = torch.device("cuda:0")
w = torch.rand(M, K, device=dev) * 2 - 1
b = torch.rand(nb, n, K, device=dev) * 2 - 1
c = torch.empty(nb, n, M, device=dev)
thr = float(torch.kthvalue(w.abs().flatten(), max(1, int(s * M * K))).values)
ri, ci, va, nnz = spfy.threshold_to_coo(w, thr)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(4):
    e0.record()
    spfy.batched.strided_coo(M, K, nnz, K, n, nb, ri, ci, va, b, c)
    e1.record()
    torch.cuda.synchronize()
    print(f"launch {i}: {e0.elapsed_time(e1)*1e3:.1f} us  nnz={nnz}")
