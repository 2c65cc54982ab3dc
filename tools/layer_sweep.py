#!/usr/bin/env python3
"""Per-shape spmma sweep over a datasets/*.csv table (unique shapes), cold operands.

For every unique (M, K, N) of the table: median-of-R CUDA-event time per launch of a train of
spfy_spmma launches that rotate over enough operand sets to stay out of L2, algorithmic GB/s and dense-equivalent TFLOP/s, and the fraction
of the per-shape roofline min(HBM, sparse tensor).  Writes a CSV line per shape to stdout.

    python tools/layer_sweep.py [--csv resnet50.csv] [--batch 32] [--dtype fp16] [--reps 7]
"""
import argparse
import collections
import json
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
    ap.add_argument("--dtype", default="fp16")
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--warm", action="store_true", help="one buffer set: operands may stay in L2")
    ap.add_argument("--train", type=int, default=16, help="launches per timed train")
    ap.add_argument("--tag", default="")
    ap.add_argument("--only", default="", help="M,K,N filter")
    ap.add_argument("--plan", action="store_true", help="also time the whole table as ONE plan (grouped launches)")
    ap.add_argument("--plan-only", action="store_true", help="skip the per-shape part")
    ap.add_argument("--plan-shapes", default="", help="semicolon list of M,K,N: plan over those layers of the table only")
    args = ap.parse_args()
    import torch
    spfy = ge.load_package()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    tdt = torch.float16 if args.dtype == "fp16" else torch.bfloat16
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(p)) if os.path.exists(p) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    hbm, tc = peaks["hbm_gbs"], 2 * peaks["bf16_tflops"]
    cnt = collections.Counter(spfy.shapes.to_gemm(s, "weights", args.batch) for s in spfy.shapes.read_shapes(args.csv))
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    print("tag,M,K,N,count,us,GBs,frac_hbm,TFLOPs,frac_tc,roofline_us,frac_roofline")
    tot_t = tot_r = 0.0
    for g, c in sorted(cnt.items(), key=lambda kv: (-kv[0].N, kv[0].M, kv[0].K)):
        by, fl = spfy.shapes.spmma_bytes(g), spfy.shapes.spmma_flops(g)
        if args.plan_only:
            tot_r += max(by / hbm / 1e3, fl / tc / 1e6) * c
            continue
        if args.only and args.only != f"{g.M},{g.K},{g.N}":
            continue
        w = (torch.rand(g.M, g.K, device=dev) * 2 - 1).to(tdt)
        comp = spfy.prune24(w)
        # rotate over enough (B, D) sets that a launch never finds its operands in the 126 MB L2, and time
        # a train of launches under one event pair (one launch is 20-90 us; the event clock ticks at ~2 us)
        per_set = (g.K * g.N + g.M * g.N) * 2
        nsets = 1 if args.warm else max(2, min(16, -(-400_000_000 // per_set)))
        b = [(torch.rand(g.K, g.N, device=dev) * 2 - 1).to(tdt) for _ in range(nsets)]
        d = [torch.empty(g.M, g.N, device=dev, dtype=tdt) for _ in range(nsets)]
        for i in range(nsets):
            spfy.spmma_compressed(comp, b[i], out=d[i])
        ts = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(args.reps):
            torch.cuda.synchronize()
            e0.record()
            for i in range(args.train):
                spfy.spmma_compressed(comp, b[i % nsets], out=d[i % nsets])
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / args.train)
        us = statistics.median(ts)
        by, fl = spfy.shapes.spmma_bytes(g), spfy.shapes.spmma_flops(g)
        roof = max(by / hbm / 1e3, fl / tc / 1e6)
        tot_t += us * c
        tot_r += roof * c
        print(f"{args.tag},{g.M},{g.K},{g.N},{c},{us:.1f},{by/us/1e3:.0f},{by/us/1e3/hbm:.3f},{fl/us/1e6:.1f},"
              f"{fl/us/1e6/tc:.3f},{roof:.1f},{roof/us:.3f}")
        del w, b, d, comp
    if tot_t:
        print(f"# {args.tag} total {tot_t:.0f} us vs roofline {tot_r:.0f} us -> {tot_r/tot_t:.3f}")
    if args.plan or args.plan_only:
        gemms = [spfy.shapes.to_gemm(s, "weights", args.batch) for s in spfy.shapes.read_shapes(args.csv)]
        if args.plan_shapes:
            keep = set(args.plan_shapes.split(";"))
            gemms = [g for g in gemms if f"{g.M},{g.K},{g.N}" in keep]
            tot_r = sum(max(spfy.shapes.spmma_bytes(g) / hbm / 1e3, spfy.shapes.spmma_flops(g) / tc / 1e6) for g in gemms)
        problems = []
        for g in gemms:
            w = (torch.rand(g.M, g.K, device=dev) * 2 - 1).to(tdt)
            problems.append(dict(comp=spfy.prune24(w), b=(torch.rand(g.K, g.N, device=dev) * 2 - 1).to(tdt),
                                 out=torch.empty(g.M, g.N, device=dev, dtype=tdt)))
        plan = spfy.SpmmaPlan(problems)
        for _ in range(3):
            plan.run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            e0.record()
            plan.run()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = statistics.median(ts)
        by = sum(spfy.shapes.spmma_bytes(g) for g in gemms)
        fl = sum(spfy.shapes.spmma_flops(g) for g in gemms)
        for i in range(plan.launches):
            tl = []
            for _ in range(args.reps):
                flush.zero_()
                e0.record()
                plan.run_launch(i)
                e1.record()
                torch.cuda.synchronize()
                tl.append(e0.elapsed_time(e1) * 1e3)
            print(f"# {args.tag} plan launch {i}: {plan.launch_info(i)} {statistics.median(tl):.0f} us")
        print(f"# {args.tag} PLAN {len(gemms)} layers in {plan.launches} launches: {us:.0f} us, {by/us/1e3:.0f} GB/s "
              f"({by/us/1e3/hbm:.3f} of HBM), {fl/us/1e6:.1f} TFLOP/s, roofline {tot_r:.0f} us -> {tot_r/us:.3f}")


if __name__ == "__main__":
    main()
