#!/usr/bin/env python3
"""One threshold_to_coo case through the C ABI with preallocated outputs (the command ncu wraps).
    python tools/thr_one.py rows cols keep_fraction [dtype f32|f16]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import torch  # noqa: E402

rows, cols = int(sys.argv[1]), int(sys.argv[2])
keep = float(sys.argv[3])
f16 = len(sys.argv) > 4 and sys.argv[4] == "f16"
spfy = ge.load_package()
from importlib import import_module  # noqa: E402
capi = import_module(spfy.__name__ + ".capi")
dev = torch.device("cuda:0")
w = torch.rand(rows, cols, device=dev) * 2 - 1
if f16:
    w = w.half()
thr = 1.0 - keep
cap = rows * cols
ri = torch.empty(cap, dtype=torch.int32, device=dev)
ci = torch.empty(cap, dtype=torch.int32, device=dev)
va = torch.empty(cap, dtype=torch.float32, device=dev)
nnz = torch.zeros(1, dtype=torch.int64, device=dev)
rp = torch.empty(rows + 1, dtype=torch.int32, device=dev)
wb = ctypes.c_size_t()
capi.spfy_threshold_workspace_bytes(rows, cols, ctypes.byref(wb))
ws = torch.empty(wb.value, dtype=torch.uint8, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s = torch.cuda.current_stream().cuda_stream
for i in range(4):
    e0.record()
    capi.spfy_threshold_to_coo(0 if f16 else 2, w.data_ptr(), cols, rows, cols, thr, ri.data_ptr(), ci.data_ptr(), va.data_ptr(), cap,
                               nnz.data_ptr(), rp.data_ptr(), ws.data_ptr(), ws.numel(), s)
    e1.record()
    torch.cuda.synchronize()
    n = int(nnz.item())
    by = rows * cols * w.element_size() + 12 * n + 4 * (rows + 1)
    us = e0.elapsed_time(e1) * 1e3
    print(f"call {i}: {us:.1f} us  nnz={n}  {by / us / 1e3:.0f} GB/s (read once + 12 B per entry)")
