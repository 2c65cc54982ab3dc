#!/usr/bin/env python3
"""Implicit GEMM vs the explicit (unfolded) operand over the 3 x 3 layers of a datasets/*.csv table.

For every unique 3 x 3 layer (k = 9 * C_in with C_in % 64 == 0; m = Ho * Wo, pad 1, stride 1 assumed -- the CSV does
not record the stride) at batch b: time
  * spfy_spmma on the materialised K x N operand (what the reference's drivers multiply; the unfold itself -- 9x the
    activation bytes written and read once more -- is NOT included),
  * spfy_spmma_conv on the NHWC activations (B never exists),
cold (L2 flushed before every launch, median of 7).  Bytes: explicit = 2KN + 2MN + 1.125MK; implicit = 2*b*H*W*C (the
activations once) + 2MN + 1.125MK.  One CSV line per layer on stdout.

    python tools/conv_sweep.py [--csv resnet50.csv] [--batch 32]
"""
import argparse
import collections
import json
import math
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--csv", default="resnet50.csv")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    import torch
    import torch.nn.functional as F
    spfy = ge.load_package()
    dev = torch.device("cuda:0")
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0
    shapes = spfy.shapes.read_shapes(args.csv)
    cnt = collections.Counter((s.m, s.n, s.k) for s in shapes if s.k % 9 == 0 and (s.k // 9) % 64 == 0
                              and math.isqrt(s.m) ** 2 == s.m)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn):
        ts = []
        for _ in range(7):
            flush.zero_()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return statistics.median(ts)

    print("tag,Ho,C_in,C_out,batch,count,explicit_us,implicit_us,speedup,explicit_frac_hbm,implicit_frac_hbm_of_its_own_bytes,"
          "implicit_GBs_algorithmic,identical")
    tot_e = tot_i = 0.0
    for (m, cout, k), c in sorted(cnt.items(), key=lambda kv: -kv[0][0]):
        ho, cin, nb = math.isqrt(m), k // 9, args.batch
        gen = torch.Generator(device=dev)
        gen.manual_seed(m + k)
        x = (torch.rand(nb, ho, ho, cin, device=dev, generator=gen) * 2 - 1).half()
        w = (torch.rand(cout, k, device=dev, generator=gen) * 2 - 1).half()
        comp = spfy.prune24(spfy.permute_conv_weights(w, cin, 3, 3))
        cols = F.unfold(x.permute(0, 3, 1, 2).float(), 3, padding=1).view(nb, cin, 9, m).permute(2, 1, 0, 3)
        b = cols.reshape(k, nb * m).half().contiguous()
        del cols
        d1 = torch.empty(cout, nb * m, dtype=torch.float16, device=dev)
        d2 = torch.empty_like(d1)
        te = timed(lambda: spfy.spmma_compressed(comp, b, out=d1))
        ti = timed(lambda: spfy.spmma_conv(comp, x, 3, 3, stride=1, pad=1, out=d2))
        same = bool(torch.equal(d1, d2))
        n = nb * m
        be = 2 * k * n + 2 * cout * n + 1.125 * cout * k
        bi = 2 * x.numel() + 2 * cout * n + 1.125 * cout * k
        print(f"{args.tag},{ho},{cin},{cout},{nb},{c},{te:.1f},{ti:.1f},{te/ti:.2f},{be/te/1e3/hbm:.3f},{bi/ti/1e3/hbm:.3f},"
              f"{bi/ti/1e3:.0f},{same}", flush=True)
        tot_e += te * c
        tot_i += ti * c
        del x, b, d1, d2
    print(f"# {args.tag} {args.csv} 3x3 layers at batch {args.batch}: explicit operand {tot_e:.0f} us, implicit GEMM {tot_i:.0f} us "
          f"-> {tot_e/tot_i:.2f}x (the explicit figure does not include producing the unfolded operand)")


if __name__ == "__main__":
    main()
