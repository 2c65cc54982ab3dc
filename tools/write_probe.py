#!/usr/bin/env python3
"""DRAM write / read / copy rates of plain torch kernels on this GPU (context for the write-heavy spmma classes)."""
import torch
dev = torch.device("cuda:0")
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device=dev)
b = torch.empty(n, dtype=torch.uint8, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def t(fn, by):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    return by / ms / 1e6
print(f"memset 1 GiB : {t(lambda: a.zero_(), n):.0f} GB/s")
print(f"copy 1 GiB   : {t(lambda: b.copy_(a), 2 * n):.0f} GB/s (read+write)")
h = a.view(torch.float16)
print(f"read (sum)   : {t(lambda: h.sum(), n):.0f} GB/s")
