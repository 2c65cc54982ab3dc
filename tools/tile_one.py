#!/usr/bin/env python3
"""One TILE_MAG prune of a large matrix (what ncu wraps): python tools/tile_one.py [rows cols]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import torch  # noqa: E402

rows, cols = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16384, 16384)
spfy = ge.load_package()
dev = torch.device("cuda:0")
w = (torch.rand(rows, cols, device=dev) * 2 - 1).half()
out = torch.empty_like(w)
for _ in range(3):
    spfy.prune24(w, out_dense=out, mode=spfy.PRUNE_TILE_MAG, compress=False)
torch.cuda.synchronize()
print("ok", spfy.prune24_check(out))
