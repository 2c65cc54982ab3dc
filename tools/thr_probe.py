#!/usr/bin/env python3
"""threshold_to_coo on ResNet-sized weight matrices: event time per call vs the kernels' own time (C ABI called
directly with preallocated outputs, no Python allocations in the loop)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
import torch
spfy = ge.load_package()
capi = spfy.capi
dev = torch.device("cuda:0")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for M, K in ((64, 576), (256, 2304), (512, 4608), (2048, 512)):
    w = torch.rand(M, K, device=dev) * 2 - 1
    cap = M * K
    ri = torch.empty(cap, dtype=torch.int32, device=dev); ci = torch.empty_like(ri)
    va = torch.empty(cap, dtype=torch.float32, device=dev)
    nnz = torch.zeros(1, dtype=torch.int64, device=dev); rp = torch.empty(M + 1, dtype=torch.int32, device=dev)
    wb = ctypes.c_size_t(); capi.spfy_threshold_workspace_bytes(M, K, ctypes.byref(wb))
    ws = torch.empty(max(wb.value, 16), dtype=torch.uint8, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    def raw():
        capi.spfy_threshold_to_coo(2, w.data_ptr(), K, M, K, 0.9, ri.data_ptr(), ci.data_ptr(), va.data_ptr(), cap,
                                   nnz.data_ptr(), rp.data_ptr(), ws.data_ptr(), ws.numel(), st)
    for fn, name in ((raw, "C ABI, preallocated"), (lambda: spfy.threshold_to_coo(w, 0.9, sync=False), "python wrapper, sync=False")):
        for _ in range(5): fn()
        torch.cuda.synchronize(); e0.record()
        for _ in range(50): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"{M}x{K} {name}: {e0.elapsed_time(e1)/50*1e3:.1f} us per call")
