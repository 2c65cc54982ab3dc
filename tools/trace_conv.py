#!/usr/bin/env python3
"""Trace of one implicit-GEMM convolution layer (dev build): SPFY_LIB=.../lib_dev/... python tools/trace_conv.py [batch ho c_in c_out out.csv]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import torch  # noqa: E402

nb, ho, cin, cout = (int(v) for v in sys.argv[1:5]) if len(sys.argv) > 4 else (32, 112, 64, 64)
out_path = sys.argv[5] if len(sys.argv) > 5 else "gpurun_out/trace_conv.csv"
spfy = ge.load_package()
dev = torch.device("cuda:0")
x = (torch.rand(nb, ho, ho, cin, device=dev) * 2 - 1).half()
w = (torch.rand(cout, 9 * cin, device=dev) * 2 - 1).half()
comp = spfy.prune24(spfy.permute_conv_weights(w, cin, 3, 3))
out = torch.empty(cout, nb * ho * ho, dtype=torch.float16, device=dev)
for _ in range(3):
    spfy.spmma_conv(comp, x, 3, 3, stride=1, pad=1, out=out)
torch.cuda.synchronize()
os.environ["SPFY_SPMMA_TRACE"] = out_path
spfy.spmma_conv(comp, x, 3, 3, stride=1, pad=1, out=out)
torch.cuda.synchronize()
del os.environ["SPFY_SPMMA_TRACE"]
print(open(out_path).read())
