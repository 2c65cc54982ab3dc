#!/usr/bin/env python3
"""Row-major against transposed (NHWC) output of spmma in steady state: the write-heavy launch classes of the ResNet-50
table as plans (several layers back to back, operands far larger than L2 -- a single flushed call hides up to 126 MB of
its stores in L2), the output as m rows 2*N bytes apart or as one contiguous run per unit; and the same with B given
K-major (opB = T: NHWC activations of a 1 x 1 convolution).
    python tools/conv_nhwc_probe.py"""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import torch  # noqa: E402

spfy = ge.load_package()
dev = torch.device("cuda:0")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, reps=7):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


CLASSES = {
    "k<=64": [(64, 64, 401408)] + [(256, 64, 401408)] * 3,
    "512x128 / 128x512": [(512, 128, 100352)] * 4 + [(128, 512, 100352)] * 3,
    "1024x256": [(1024, 256, 25088)] * 6,
    "2048x512 / 256x512": [(256, 512, 100352)] + [(2048, 512, 6272)] * 3,
    "64x576 / 64x256 (read-heavy)": [(64, 576, 401408)] * 3 + [(64, 256, 401408)] * 2,
}
print("class,layers,MB,rowmajor_us,nhwc_us,bt_nhwc_us,rowmajor_TBs,nhwc_TBs,bt_nhwc_TBs")
for name, shapes in CLASSES.items():
    res, mb = [], 0.0
    for mode in ("row", "nhwc", "bt_nhwc"):
        probs = []
        for (M, K, N) in shapes:
            comp = spfy.prune24((torch.rand(M, K, device=dev) * 2 - 1).half())
            b = (torch.rand(N, K, device=dev) * 2 - 1).half() if mode == "bt_nhwc" else (torch.rand(K, N, device=dev) * 2 - 1).half()
            out = torch.empty((M, N) if mode == "row" else (N, M), dtype=torch.float16, device=dev)
            probs.append(dict(comp=comp, b=b, out=out, out_t=mode != "row", op_b=spfy.OP_T if mode == "bt_nhwc" else spfy.OP_N))
        plan = spfy.SpmmaPlan(probs)
        res.append(timed(plan.run))
        plan.close()
        mb = sum((K * N + M * N) * 2 + M * K * 1.125 for (M, K, N) in shapes) / 1e6
        del probs
    print(f"{name},{len(shapes)},{mb:.0f},{res[0]:.1f},{res[1]:.1f},{res[2]:.1f},{mb/res[0]:.2f},{mb/res[1]:.2f},{mb/res[2]:.2f}", flush=True)
