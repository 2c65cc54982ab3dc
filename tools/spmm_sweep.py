#!/usr/bin/env python3
"""Unstructured path over a datasets/*.csv table: magnitude threshold -> COO, then batched COO SpMM.

BASELINE.json configs[2] / configs[3]: for every unique shape of the table (weights orientation:
A = W [C_out x K] shared by the batch, B_b = [K x H*W] column-major per image, C_b = [C_out x H*W]
column-major, fp32 -- the operand types of spmm.hxx:165-180) and every sparsity s in --sparsity, time
  * spfy_threshold_to_coo    (threshold = the ceil(s*M*K)-th smallest |w|),
  * spfy_spmm_coo_strided_batched over nb = b images,
one CUDA-event pair per call, median of `--reps` calls (operands of one launch exceed L2 for every
ResNet shape at b=32).  Algorithmic bytes: threshold 4*M*K*2 reads + 12*nnz;
SpMM 12*nnz + 4*K*N*nb + 4*M*N*nb.  One CSV line per (shape, sparsity) on stdout.

    python tools/spmm_sweep.py [--csv resnet34.csv] [--batch 32] [--sparsity 0.5,0.9,0.95]
"""
import argparse
import collections
import statistics
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--csv", default="resnet34.csv")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--sparsity", default="0.5,0.9,0.95")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--tag", default="")
    ap.add_argument("--no-cusparse", action="store_true", help="skip the cuSPARSE comparator column")
    ap.add_argument("--alg", default="DEFAULT", choices=["DEFAULT", "CUDA_CORE", "TENSOR", "TENSOR_FAST"],
                    help="SPFY_SPMM_ALG_* of include/spfy_b200.h (DEFAULT: tcgen05 3xTF32 when A is dense enough to pay)")
    args = ap.parse_args()
    import torch
    spfy = ge.load_package()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0
    shapes = spfy.shapes.read_shapes(args.csv)
    cnt = collections.Counter((s.n, s.k, s.m) for s in shapes)  # (M, K, n = H*W)
    nb = args.batch
    alg = getattr(spfy, "SPMM_ALG_" + args.alg)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cusp = os.path.join(ROOT, "oracle", "_ref", "cusparse_ref")  # comparator: cuSPARSE COO_ALG4 on the same GPU
    have_cusp = os.path.exists(cusp) and not args.no_cusparse
    print("tag,sparsity,M,K,n,nb,count,nnz,thr_us,thr_GBs,spmm_us,spmm_GBs,spmm_frac_hbm,spmm_GFLOPs,cusparse_us")
    tot_c = collections.defaultdict(float)
    tot = collections.defaultdict(lambda: [0.0, 0.0])
    for (M, K, n), c in sorted(cnt.items(), key=lambda kv: (-kv[0][2], kv[0][0], kv[0][1])):
        gen = torch.Generator(device=dev)
        gen.manual_seed(0x5EED)
        w = torch.rand(M, K, device=dev, generator=gen) * 2 - 1
        b = torch.rand(nb, n, K, device=dev, generator=gen) * 2 - 1   # nb slabs of [K x n] column-major
        cbuf = torch.empty(nb, n, M, device=dev)                        # nb slabs of [M x n] column-major
        for s in (float(x) for x in args.sparsity.split(",")):
            kth = max(1, min(M * K, int(-(-s * M * K // 1))))
            thr = float(torch.kthvalue(w.abs().flatten(), kth).values)
            ri, ci, va, nnz = spfy.threshold_to_coo(w, thr)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.reps):
                spfy.threshold_to_coo(w, thr)
            e1.record()
            torch.cuda.synchronize()
            thr_us = e0.elapsed_time(e1) / args.reps * 1e3
            try:
                spfy.batched.strided_coo(M, K, nnz, K, n, nb, ri, ci, va, b, cbuf, alg=alg)
            except spfy.SpfyError:  # forced tensor route on an operand TMA cannot address (k = 147)
                alg_here = spfy.SPMM_ALG_DEFAULT
            else:
                alg_here = alg
            torch.cuda.synchronize()
            # the wrapper brackets the call with an event pair and synchronises (the reference's timer_t semantics): its
            # own figure is the time of the call on the GPU; a loop under one outer pair would add a host round trip each
            us = statistics.median(spfy.batched.strided_coo(M, K, nnz, K, n, nb, ri, ci, va, b, cbuf, alg=alg_here)
                                   for _ in range(args.reps)) * 1e3
            thr_bytes = 2 * 4 * M * K + 12 * nnz
            by = 12 * nnz + 4 * K * n * nb + 4 * M * n * nb
            fl = 2.0 * nnz * n * nb
            cus = ""
            if have_cusp:
                try:
                    torch.cuda.synchronize()
                    r = subprocess.run([cusp, "time", str(M), str(K), str(n), str(nb), str(1.0 - s)], capture_output=True,
                                       text=True, timeout=300)
                    cus = json.loads(r.stdout.strip().splitlines()[-1])["us"]
                    tot_c[s] += cus * c
                except Exception:  # noqa: BLE001
                    cus = ""
            print(f"{args.tag},{s},{M},{K},{n},{nb},{c},{nnz},{thr_us:.1f},{thr_bytes/thr_us/1e3:.0f},{us:.1f},"
                  f"{by/us/1e3:.0f},{by/us/1e3/hbm:.3f},{fl/us/1e3:.0f},{cus}", flush=True)
            tot[s][0] += us * c
            tot[s][1] += by / (hbm * 1e3) * c
        del w, b, cbuf
    for s, (t, r) in tot.items():
        extra = f"; cuSPARSE COO_ALG4 {tot_c[s]:.0f} us" if tot_c.get(s) else ""
        print(f"# {args.tag} sparsity {s}: table total {t:.0f} us vs HBM roofline {r:.0f} us -> {r/t:.3f}{extra}")


if __name__ == "__main__":
    main()
