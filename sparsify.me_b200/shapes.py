"""Shape tables of the per-layer conv-as-GEMM problems (datasets/*.csv).

Mirrors util::read_shapes (reference: include/sparsify.me/util/util.hxx:36-61): skip the header,
split on commas, integer fields (m, n, k, b); CRLF line endings are tolerated.
CSV columns follow datasets/get_shapes.py:68-74 of the reference:
    m = H_out*W_out, n = C_out, k = C_in*kh*kw, b = image batch
Orientations:
    "ref"     : A is m x k (what the reference drivers pass: examples/profiling.py:39-41), N = n
    "weights" : A is the weight matrix, M = n_csv, K = k_csv, N = m_csv * b   (the north star)
"""
import os
from collections import namedtuple

DATASETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "datasets")

Shape = namedtuple("Shape", "m n k b")
Gemm = namedtuple("Gemm", "M N K")


def read_shapes(path):
    if not os.path.isabs(path) and not os.path.exists(path):
        path = os.path.join(DATASETS, path)
    out = []
    with open(path, "r", newline="") as f:
        lines = f.read().split("\n")
    for line in lines[1:]:
        line = line.strip("\r ")
        if not line:
            continue
        fields = line.split(",")
        if len(fields) != 4:
            raise ValueError("read_shapes: expected 4 comma-separated fields, got %r" % (line,))
        out.append(Shape(*(int(x) for x in fields)))
    return out


def to_gemm(shape, orient="weights", batch=None):
    b = shape.b if batch is None else batch
    if orient == "weights":
        return Gemm(M=shape.n, N=shape.m * b, K=shape.k)
    if orient == "ref":
        return Gemm(M=shape.m, N=shape.n, K=shape.k)
    raise ValueError("orient must be 'weights' or 'ref'")


def pad8(x):
    return (x + 7) // 8 * 8


def spmma_flops(g):
    """dense-equivalent FLOPs, 2*M*N*K"""
    return 2.0 * g.M * g.N * g.K


def spmma_bytes(g, elem=2, beta=False):
    """algorithmic HBM bytes of one spmma: B + D (+C) + compressed A values + metadata"""
    return elem * g.K * g.N + elem * g.M * g.N * (2 if beta else 1) + g.M * g.K * elem // 2 + g.M * g.K // 8


def prune24_bytes(rows, cols, elem=2, dense_out=False, mask=False):
    """algorithmic HBM bytes of prune+compress: read + values + metadata (3.125 B/elem at 16 bit)"""
    e = rows * cols
    return e * elem + e * elem // 2 + e // 8 + (e * elem if dense_out else 0) + (8 * e if mask else 0)
