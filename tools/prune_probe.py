#!/usr/bin/env python3
"""prune24 (2:4 magnitude prune + compress + metadata) bandwidth on matrices large enough to be
HBM-bound rather than launch-bound.

The per-layer weight matrices of datasets/*.csv are <= 4.7 MB (a per-layer prune is
launch-latency-bound), so the kernel's bandwidth is measured here on
  * the `ref`-orientation operand of the reference drivers (examples/profiling.py:39-41 passes the
    m x k activation-shaped matrix: 12544*32 x 576 fp16 = 462 MB),
  * a square 16384 x 16384 matrix (512 MB),
  * and, for scale, the whole ResNet-50 weight set in one batched launch (47 MB).
Algorithmic bytes: 3.125 B/element (read 2, values 1, metadata 1/8), +2 with the pruned dense copy.

    python tools/prune_probe.py [--reps 10] [--dtype fp16]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--dtype", default="fp16")
    ap.add_argument("--tag", default="")
    ap.add_argument("--only-large", action="store_true", help="just the 401408 x 576 SM100 case (what ncu wraps)")
    ap.add_argument("--tile", action="store_true", help="TILE_MAG (the mode spmma.hxx:86 requests): dense prune alone, and prune + compress")
    args = ap.parse_args()
    import torch
    spfy = ge.load_package()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    tdt = torch.float16 if args.dtype == "fp16" else torch.bfloat16
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.reps * 1e3

    print("tag,case,rows,cols,layout,dense_out,us,GBs,frac_hbm")
    if args.tile:
        # TILE needs the dense pruned matrix as an output: 2 B read + 2 B written per element (+ 1.125 compressed)
        for rows, cols in ((16384, 16384), (12544 * 32, 576), (4096, 4608)):
            w = (torch.rand(rows, cols, device=dev) * 2 - 1).to(tdt)
            out = torch.empty_like(w)
            us = timed(lambda: spfy.prune24(w, out_dense=out, mode=spfy.PRUNE_TILE_MAG, compress=False))
            by = 4.0 * rows * cols
            print(f"{args.tag},tile-dense,{rows},{cols},-,1,{us:.1f},{by/us/1e3:.0f},{by/us/1e3/hbm:.3f}", flush=True)
            comp = spfy.alloc_compressed(tdt, rows, cols, dev)
            us = timed(lambda: spfy.prune24(w, out_dense=out, mode=spfy.PRUNE_TILE_MAG, out=comp))
            by = 5.125 * rows * cols
            print(f"{args.tag},tile-compress,{rows},{cols},sm100,1,{us:.1f},{by/us/1e3:.0f},{by/us/1e3/hbm:.3f}", flush=True)
            del w, out, comp
        return
    cases = ((12544 * 32, 576),) if args.only_large else ((12544 * 32, 576), (16384, 16384), (4096, 4608))
    for rows, cols in cases:
        w = (torch.rand(rows, cols, device=dev) * 2 - 1).to(tdt)
        for layout, lname in ((spfy.LAYOUT_SM100, "sm100"), (spfy.LAYOUT_CANONICAL, "canonical"))[:1 if args.only_large else 2]:
            comp = spfy.alloc_compressed(tdt, rows, cols, dev, layout)
            for dense in (False, True):
                out = torch.empty_like(w) if dense else None
                us = timed(lambda: spfy.prune24(w, out_dense=out, layout=layout, out=comp))
                by = spfy.shapes.prune24_bytes(rows, cols, dense_out=dense)
                print(f"{args.tag},single,{rows},{cols},{lname},{int(dense)},{us:.1f},{by/us/1e3:.0f},{by/us/1e3/hbm:.3f}",
                      flush=True)
                del out
            del comp
        del w
    if args.only_large:
        return
    gemms = [spfy.shapes.to_gemm(s, "weights", 32) for s in spfy.shapes.read_shapes("resnet50.csv")]
    ws = [(torch.rand(g.M, g.K, device=dev) * 2 - 1).to(tdt) for g in gemms]
    comps = [spfy.alloc_compressed(tdt, g.M, g.K, dev) for g in gemms]
    us = timed(lambda: spfy.prune24_batched(ws, comps))
    by = sum(spfy.shapes.prune24_bytes(g.M, g.K) for g in gemms)
    print(f"{args.tag},resnet50-batched,{len(gemms)},-,sm100,0,{us:.1f},{by/us/1e3:.0f},{by/us/1e3/hbm:.3f}")


if __name__ == "__main__":
    main()
