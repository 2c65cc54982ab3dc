#!/usr/bin/env python3
"""Probe of the dense tcgen05 GEMM (spfy_gemm_strided_batched): one line per (dtype, opA, opB, shape) with the
worst error relative to sum|a||b| against fp64, and optional timing against torch.matmul (cuBLAS).
    python tools/gemm_probe.py [--dtype f32|f16|bf16] [--time]
A configuration that faults takes the CUDA context with it: each line is printed (flushed) BEFORE its launch."""
import argparse
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--time", action="store_true")
    args = ap.parse_args()
    import torch
    spfy = ge.load_package()
    dev = torch.device("cuda:0")
    tdt = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}[args.dtype]
    shapes = [(128, 128, 64, 1), (128, 64, 32, 1), (256, 256, 256, 2), (200, 72, 104, 3), (64, 392, 256, 2)]
    if not args.time:
        for ta, tb in [(1, 0), (0, 0), (1, 1), (0, 1)]:
            for prec in ([0, 1] if args.dtype == "f32" else [0]):
                for (m, n, k, nb) in shapes:
                    print(f"probe {args.dtype} ta={ta} tb={tb} prec={prec} m={m} n={n} k={k} nb={nb} ...", end=" ", flush=True)
                    g = torch.Generator(device=dev)
                    g.manual_seed(m + n + k)
                    A = (torch.rand(nb, m, k, device=dev, generator=g) * 2 - 1).to(tdt)
                    B = (torch.rand(nb, k, n, device=dev, generator=g) * 2 - 1).to(tdt)
                    sa = A.transpose(1, 2).contiguous() if ta == 0 else A.contiguous()  # N: column-major m x k = [k, m]
                    sb = B.transpose(1, 2).contiguous() if tb == 0 else B.contiguous()
                    C = torch.full((nb, n, m), 7.0, device=dev, dtype=tdt)
                    try:
                        spfy.batched.gemm(sa, sb, C, m, n, k, transpose_a=ta, transpose_b=tb, precision=prec)
                        torch.cuda.synchronize()
                    except Exception as e:  # noqa: BLE001
                        print("ERROR", str(e)[:200], flush=True)
                        continue
                    want = A.double() @ B.double()
                    bound = A.double().abs() @ B.double().abs()
                    got = C.double().transpose(1, 2)
                    err = float(((got - want).abs() / bound).max())
                    print(f"max err / sum|a||b| = {err:.3e}", flush=True)
        return
    # timing: the three shapes that matter (K-major both: the SpMM routes; MN-major A: batched::gemm of the drivers)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for (m, n, k, nb, ta) in [(12544, 64, 576, 32, 1), (3136, 128, 1152, 32, 1), (784, 256, 2304, 32, 1), (196, 512, 4608, 32, 1),
                              (12544, 64, 576, 32, 0), (784, 256, 2304, 32, 0)]:
        A = torch.rand(nb, m, k, device=dev).to(tdt) if ta else torch.rand(nb, k, m, device=dev).to(tdt)
        B = torch.rand(n, k, device=dev).to(tdt)
        C = torch.empty(nb, n, m, device=dev, dtype=tdt)
        for prec in ([0, 1] if args.dtype == "f32" else [0]):
            for _ in range(2):
                spfy.batched.gemm(A, B, C, m, n, k, transpose_a=ta, precision=prec)
            torch.cuda.synchronize()
            # the wrapper brackets the call with an event pair and synchronises (like the reference's timer_t), so a
            # loop under ONE outer event pair would time the host round trips, not the kernel: take its own figure
            us = statistics.median(spfy.batched.gemm(A, B, C, m, n, k, transpose_a=ta, precision=prec) for _ in range(9)) * 1e3
            fl = 2.0 * m * n * k * nb
            by = (m * k * nb + n * k + m * n * nb) * A.element_size()
            print(f"time {args.dtype} ta={ta} prec={prec} m={m} n={n} k={k} nb={nb}: {us:.1f} us  {fl/us/1e6:.1f} TFLOP/s  "
                  f"{by/us/1e3:.0f} GB/s", flush=True)
        if ta:
            Bt = B.t().contiguous()
            for _ in range(2):
                torch.matmul(A, Bt)
            torch.cuda.synchronize()
            ts = []
            for _ in range(9):  # same method: one event pair per call
                e0.record()
                torch.matmul(A, Bt)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            us = statistics.median(ts)
            print(f"     torch.matmul (cuBLAS, {'fp32 no-TF32' if tdt == torch.float32 else args.dtype}): {us:.1f} us", flush=True)


if __name__ == "__main__":
    main()
