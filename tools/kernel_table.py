#!/usr/bin/env python3
"""One line per kernel of the path: time, algorithmic bytes, GB/s and the
fraction of the measured HBM copy rate, each on an input large enough to be bandwidth- rather than
launch-bound.  CSV on stdout (kept as profiles/rNN_kernels.csv).

    python tools/kernel_table.py [--reps 10]
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
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    import torch
    spfy = ge.load_package()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, reps=args.reps):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3

    print("tag,row,kernel,case,us,algorithmic_MB,GBs,frac_hbm,note")

    def line(row, kernel, case, us, by, note=""):
        print(f"{args.tag},{row},{kernel},{case},{us:.1f},{by/1e6:.1f},{by/us/1e3:.0f},{by/us/1e3/hbm:.3f},{note}", flush=True)

    # ---- A1 positional sparsify<2,2> (sparsify.hxx:24-82): the driver's fp32 m x k operand, u64 mask
    m, n = 12544 * 32, 147
    w = torch.rand(m * n, device=dev)
    mask = torch.empty(m * n, dtype=torch.int64, device=dev)
    us = timed(lambda: spfy.sparsify(w, mask, m, n))
    line("A1", "prune_blocks_ref_kernel", f"fp32 {m}x{n} blk 2x2", us, m * n * 8 + m * n * 4 // 2,
         "write-only: 8 B mask per element + the zeroed half of the weights")
    del w, mask

    # ---- A2/A3 prune24 (spmma.hxx:85-104)
    rows, cols = 16384, 16384
    a = (torch.rand(rows, cols, device=dev) * 2 - 1).half()
    for layout, name in ((spfy.LAYOUT_SM100, "sm100"), (spfy.LAYOUT_CANONICAL, "canonical")):
        comp = spfy.alloc_compressed(torch.float16, rows, cols, dev, layout)
        us = timed(lambda: spfy.prune24(a, layout=layout, out=comp))
        line("A2/A3", "prune24_fast_kernel", f"fp16 {rows}x{cols} -> {name}", us, spfy.shapes.prune24_bytes(rows, cols),
             "read 2 + values 1 + metadata 1/8 B per element")
        del comp
    dense = torch.empty_like(a)
    us = timed(lambda: spfy.prune24(a, out_dense=dense, compress=False))
    line("A2", "prune24_fast_kernel", f"fp16 {rows}x{cols} dense out only", us, rows * cols * 4, "read 2 + write 2")
    us = timed(lambda: spfy.prune24_check(dense), reps=3)
    line("A2", "prune24_check_kernel", f"fp16 {rows}x{cols}", us, rows * cols * 2, "includes the D2H of the flag (spmma.hxx:90-92)")
    dt = torch.empty_like(a)
    us = timed(lambda: spfy.prune24(a, out_dense=dt, compress=False, mode=spfy.PRUNE_TILE_MAG))
    line("A2", "prune24_tile_kernel", f"fp16 {rows}x{cols} 4x4 tiles", us, rows * cols * 4,
         "cusparseLt's TILE selection (19 candidates per tile); read 2 + write 2 B per element")
    comp = spfy.alloc_compressed(torch.float16, rows, cols, dev, spfy.LAYOUT_SM100)
    us = timed(lambda: spfy.prune24(a, out_dense=dt, mode=spfy.PRUNE_TILE_MAG, out=comp))
    line("A2/A3", "prune24_tile_kernel + prune24_fast_kernel", f"fp16 {rows}x{cols} TILE -> sm100", us,
         spfy.shapes.prune24_bytes(rows, cols), "what sparsifyme::spmma runs: two passes, algorithmic bytes of one")
    del a, dense, dt, comp
    rows_w, cols_w = 512, 4608  # the largest ResNet weight matrix: TILE prune + compress is one fused launch
    aw = (torch.rand(rows_w, cols_w, device=dev) * 2 - 1).half()
    dw = torch.empty_like(aw)
    cw = spfy.alloc_compressed(torch.float16, rows_w, cols_w, dev, spfy.LAYOUT_SM100)
    us = timed(lambda: spfy.prune24(aw, out_dense=dw, mode=spfy.PRUNE_TILE_MAG, out=cw))
    line("A2/A3", "prune24_tile_fused_kernel", f"fp16 {rows_w}x{cols_w} TILE -> sm100", us, spfy.shapes.prune24_bytes(rows_w, cols_w) + rows_w * cols_w * 2,
         "one launch: pruned dense + compressed values + metadata (launch-bound at this size)")
    us = timed(lambda: spfy.prune24(aw, out_dense=dw, out=cw))
    line("A2/A3", "prune24_fast_kernel", f"fp16 {rows_w}x{cols_w} STRIP -> sm100", us, spfy.shapes.prune24_bytes(rows_w, cols_w) + rows_w * cols_w * 2,
         "same outputs with the per-row selection")
    del aw, dw, cw

    # ---- threshold -> COO (the <todo> of sparsify.hxx:58-59)
    rows, cols = 8192, 8192
    wf = torch.rand(rows, cols, device=dev) * 2 - 1
    for s in (0.5, 0.9):
        thr = s
        ri, ci, va, nnz = spfy.threshold_to_coo(wf, thr)
        us = timed(lambda: spfy.threshold_to_coo(wf, thr), reps=5)
        line("A6-prep", "threshold_compact_kernel", f"fp32 {rows}x{cols} keep |x|>{thr}", us, 4 * rows * cols + 12 * nnz + 4 * rows,
             f"nnz={nnz}; input read once + 12 B per kept entry + row_ptr; python wrapper incl. the nnz read-back (kernel alone: tools/thr_one.py)")
    del wf

    # ---- A6 batched COO SpMM (spmm.hxx:140-193): one ResNet-34 layer per regime
    for (M, K, n, nb, s) in ((64, 576, 12544, 32, 0.95), (64, 576, 12544, 32, 0.9), (256, 2304, 784, 32, 0.9),
                             (512, 4608, 196, 32, 0.5)):
        w = torch.rand(M, K, device=dev) * 2 - 1
        b = torch.rand(nb, n, K, device=dev) * 2 - 1
        c = torch.empty(nb, n, M, device=dev)
        thr = float(torch.kthvalue(w.abs().flatten(), max(1, int(s * M * K))).values)
        ri, ci, va, nnz = spfy.threshold_to_coo(w, thr)
        us = timed(lambda: spfy.batched.strided_coo(M, K, nnz, K, n, nb, ri, ci, va, b, c), reps=5)
        by = 12 * nnz + 4 * K * n * nb + 4 * M * n * nb
        lds_us = nnz * n * nb / 32 / 148 / 1.9e3  # one shared-memory wavefront per 32 FMAs, 148 SMs, ~1.9 GHz
        line("A6", "spmm_dense_walk_kernel" if nnz >= 0.35 * M * K else "spmm_csr_kernel", f"M={M} K={K} n={n} nb={nb} sparsity {s}", us, by,
             f"nnz={nnz}; {2.0*nnz*n*nb/us/1e6:.1f} TFLOP/s fp32; shared-memory-wavefront bound {lds_us:.0f} us")
        del w, b, c

    # ---- A5 blocked-ELL batched SpMM (spmm.hxx:30-138), the reference driver's construction (block 2, ell_cols = k/2)
    m, n, k, nb, block = 512, 1024, 1024, 8, 2
    ell_cols = k // 2
    bcols = ell_cols // block
    B = torch.rand(n, k, device=dev)
    cis = [torch.stack([torch.randperm(k // block, device=dev)[:bcols].sort().values for _ in range(m // block)]).to(torch.int64)
           for _ in range(nb)]
    vas = [torch.rand(m, ell_cols, device=dev) for _ in range(nb)]
    cs = [torch.empty(n, m, device=dev) for _ in range(nb)]
    us = timed(lambda: spfy.batched.spmm(cis, vas, B, cs, m, n, k, block, ell_cols), reps=3)
    by = nb * (m * ell_cols * 4 + (m // block) * bcols * 8 + m * n * 4) + k * n * 4
    line("A5", "spmm_csr_kernel<blocked-ELL pairs>", f"m={m} n={n} k={k} nb={nb} block 2", us, by,
         f"{2.0*nb*m*ell_cols*n/us/1e6:.1f} TFLOP/s fp32")


if __name__ == "__main__":
    main()
