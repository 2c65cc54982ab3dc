#!/usr/bin/env python3
"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE on a B200.

  ref_sparsify_*.npz   outputs of the reference's own sparsifyme::sparsify<BLK_M,BLK_N>
                       (compiled from /root/reference/include/sparsify.me/sparsify.hxx by
                       oracle/Makefile -> oracle/_ref/ref_sparsify_dump) on the deterministic input
                       weights[i] = 1 + i % 251.  Small cases are stored whole, large ones as SHA-256
                       of the weight and mask buffers.
  cusparselt_*.npz     what the closed library behind the reference's spmma does
                       (oracle/_ref/cusparselt_ref golden ...: cusparseLtSpMMAPrune STRIP / TILE,
                       Compress, Matmul -- the call sequence of include/sparsify.me/spmma.hxx:51-114)
                       on inputs from the splitmix64 generator `gen` below.

  cusparse_coo_*.npz / cusparse_bell_*.npz
                       what cuSPARSE -- the library the reference delegates the unstructured SpMM to --
                       returns for the call sequences of include/sparsify.me/spmm.hxx:164-187 (batched COO,
                       CUSPARSE_SPMM_COO_ALG4) and :57-67,107-110 (blocked-ELL), driven by
                       oracle/_ref/cusparse_ref on inputs from `gen` (exact in fp32, so comparisons are
                       bit-exact).

Needs a GPU:   gpurun -- python tests/golden/make_golden.py gpurun_out/golden
then copy gpurun_out/golden/*.npz here.  The CPU test-suite (tests/test_golden.py) only READS them.
"""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref")

SPARSIFY_CASES = [  # (m, n, blk, sparsity_factor, store_whole)
    (8, 8, 22, 0.5, True), (7, 5, 22, 0.5, True), (12, 10, 22, 0.25, True), (12, 10, 22, 0.75, True),
    (16, 12, 44, 0.5, True), (16, 12, 42, 0.5, True), (9, 7, 11, 0.5, True), (9, 7, 11, 1.0, True),
    (36, 52, 22, 0.5, True), (12544, 147, 22, 0.5, False), (196, 4608, 22, 0.5, False), (3136, 576, 22, 0.5, False),
]
CUSPARSELT_CASES = [(64, 128, 64, "strip"), (64, 128, 64, "tile"), (128, 256, 136, "strip"), (128, 256, 136, "tile"),
                    (256, 576, 64, "strip")]


CUSPARSE_COO_CASES = [  # (m, k, n, nb, threshold, alpha, beta)
    (64, 147, 96, 2, 0.5, 1.0, 0.0), (128, 576, 40, 3, 0.875, 1.0, 0.0), (96, 448, 150, 2, 0.75, 0.5, 2.0),
    (33, 70, 17, 1, 0.25, 1.0, 0.0)]
CUSPARSE_BELL_CASES = [(64, 128, 48, 3, 2), (128, 448, 40, 2, 4), (64, 256, 24, 2, 16)]  # (m, k, n, nb, block)


def gen_f32(t, n):
    """`gen` as float32 (the unstructured path is fp32; the values are multiples of 1/64 either way)"""
    return gen(t, n).astype(np.float32)


def gen(t, n):
    """same splitmix64 counter generator as oracle/cusparselt_ref.cu: multiples of 1/64 in [-1, 1)"""
    i = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (np.uint64(t) * np.uint64(0x632BE59BD9B4E019) + i + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    q = (z >> np.uint64(57)).astype(np.int64) - 64
    return (q.astype(np.float32) / 64.0).astype(np.float16)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.dirname(os.path.abspath(__file__))
    os.makedirs(out, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        for m, n, blk, sf, whole in SPARSIFY_CASES:
            f = os.path.join(tmp, "s.bin")
            subprocess.run([os.path.join(REF, "ref_sparsify_dump"), str(m), str(n), str(blk), str(sf), f], check=True)
            raw = np.fromfile(f, dtype=np.uint8)
            w = raw[: m * n * 4].view(np.float32)
            mask = raw[m * n * 4:].view(np.uint64)
            name = f"ref_sparsify_{m}x{n}_blk{blk}_sf{sf}.npz"
            if whole:
                np.savez_compressed(os.path.join(out, name), m=m, n=n, blk=blk, sf=sf, weights=w, mask=mask)
            else:
                np.savez_compressed(os.path.join(out, name), m=m, n=n, blk=blk, sf=sf, weights_sha256=sha(w),
                                    mask_sha256=sha(mask), zeros=int((w == 0).sum()), mask_sum=int(mask.sum()))
            print("wrote", name)
        for m, k, n, alg in CUSPARSELT_CASES:
            f = os.path.join(tmp, "c.bin")
            r = subprocess.run([os.path.join(REF, "cusparselt_ref"), "golden", str(m), str(k), str(n), alg, f],
                               capture_output=True, text=True)
            print(r.stdout.strip(), r.stderr.strip())
            if r.returncode != 0:
                continue
            raw = np.fromfile(f, dtype=np.uint8)
            hdr = raw[:40].view(np.int64)
            body = raw[40:].view(np.uint16)
            a_in, a_pr, b, d = np.split(body, [m * k, 2 * m * k, 2 * m * k + k * n])
            assert np.array_equal(a_in, gen(1, m * k).view(np.uint16)), "generator mismatch (A)"
            assert np.array_equal(b, gen(2, k * n).view(np.uint16)), "generator mismatch (B)"
            name = f"cusparselt_{alg}_{m}x{k}x{n}.npz"
            np.savez_compressed(os.path.join(out, name), m=m, k=k, n=n, alg=alg, valid=int(hdr[4]),
                                a_pruned=a_pr.reshape(m, k), d=d.reshape(m, n))
            print("wrote", name)
        exe = os.path.join(REF, "cusparse_ref")
        for m, k, n, nb, thr, alpha, beta in CUSPARSE_COO_CASES:
            f = os.path.join(tmp, "coo.bin")
            r = subprocess.run([exe, "coo", str(m), str(k), str(n), str(nb), str(thr), str(alpha), str(beta), f],
                               capture_output=True, text=True)
            print(r.stdout.strip(), r.stderr.strip())
            if r.returncode != 0:
                continue
            raw = np.fromfile(f, dtype=np.uint8)
            nnz = int(raw[:40].view(np.int64)[4])
            body = raw[40:]
            rows = body[: 4 * nnz].view(np.int32)
            cols = body[4 * nnz: 8 * nnz].view(np.int32)
            vals = body[8 * nnz: 12 * nnz].view(np.float32)
            c = body[12 * nnz:].view(np.float32).reshape(nb, n, m)
            a = gen_f32(1, m * k).reshape(m, k)
            keep = np.abs(a) > np.float32(thr)
            assert np.array_equal(np.argwhere(keep)[:, 0].astype(np.int32), rows), "generator mismatch (COO rows)"
            assert np.array_equal(a[keep], vals), "generator mismatch (COO values)"
            name = f"cusparse_coo_{m}x{k}x{n}x{nb}.npz"
            np.savez_compressed(os.path.join(out, name), m=m, k=k, n=n, nb=nb, thr=thr, alpha=alpha, beta=beta, nnz=nnz,
                                coo_sha256=sha(np.concatenate([rows, cols])), c=c)
            print("wrote", name)
        for m, k, n, nb, block in CUSPARSE_BELL_CASES:
            f = os.path.join(tmp, "bell.bin")
            r = subprocess.run([exe, "bell", str(m), str(k), str(n), str(nb), str(block), f], capture_output=True, text=True)
            print(r.stdout.strip(), r.stderr.strip())
            if r.returncode != 0:
                continue
            raw = np.fromfile(f, dtype=np.uint8)
            ell_cols = int(raw[:48].view(np.int64)[5])
            nci = nb * (m // block) * (ell_cols // block)
            ci = raw[48: 48 + 8 * nci].view(np.int64).reshape(nb, m // block, ell_cols // block)
            c = raw[48 + 8 * nci:].view(np.float32).reshape(nb, n, m)
            name = f"cusparse_bell_{m}x{k}x{n}x{nb}_b{block}.npz"
            np.savez_compressed(os.path.join(out, name), m=m, k=k, n=n, nb=nb, block=block, ell_cols=ell_cols,
                                col_idx=ci, c=c)
            print("wrote", name)


if __name__ == "__main__":
    main()
