"""ctypes/numpy front end of oracle/liboracle.so (the CPU restatement in spfy_oracle.cpp).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package.  Arrays are numpy;
16-bit floats travel as uint16 bit patterns (fp16 / bf16) so that nothing is rounded on the way.
"""
import ctypes
import os
import subprocess
from ctypes import c_double, c_float, c_int, c_longlong, c_size_t, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
F16, BF16, F32, F64 = 0, 1, 2, 3


def build(force=False):
    src = os.path.join(HERE, "spfy_oracle.cpp")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.orc_spmma_f64.argtypes = [c_int, c_int, c_size_t, c_size_t, c_size_t, c_double, c_void_p,
                                       c_size_t, c_void_p, c_size_t, c_double, c_void_p, c_size_t,
                                       c_void_p]
        _lib.orc_spmma_compressed_f32.argtypes = [c_int, c_size_t, c_size_t, c_size_t, c_float, c_void_p,
                                                  c_void_p, c_void_p, c_size_t, c_float, c_void_p,
                                                  c_size_t, c_void_p, c_size_t]
        _lib.orc_prune_blocks_ref.argtypes = [c_int, c_void_p, c_void_p, c_size_t, c_size_t, c_size_t,
                                              c_size_t, c_float]
        _lib.orc_prune24_strip.argtypes = [c_int, c_void_p, c_size_t, c_size_t, c_size_t, c_void_p,
                                           c_size_t, c_void_p, c_void_p, c_void_p]
        _lib.orc_prune24_tile.argtypes = [c_int, c_void_p, c_size_t, c_size_t, c_size_t, c_void_p,
                                          c_size_t, c_void_p]
        _lib.orc_prune24_check.argtypes = [c_int, c_void_p, c_size_t, c_size_t, c_size_t]
        _lib.orc_pack_sm100.argtypes = [c_void_p, c_void_p, c_size_t, c_size_t, c_void_p, c_void_p]
        _lib.orc_threshold_to_coo.argtypes = [c_int, c_void_p, c_size_t, c_size_t, c_size_t, c_float,
                                              c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
        _lib.orc_threshold_to_coo.restype = c_longlong
        _lib.orc_coo_to_csr.argtypes = [c_void_p, c_size_t, c_size_t, c_void_p]
        _lib.orc_spmm_coo_batched_f64.argtypes = [c_size_t] * 5 + [c_void_p] * 4 + [c_size_t, c_size_t,
                                                                                  c_void_p, c_size_t,
                                                                                  c_size_t, c_double,
                                                                                  c_double, c_void_p]
        _lib.orc_spmm_coo_batched_f32.argtypes = [c_size_t] * 5 + [c_void_p] * 4 + [c_size_t, c_size_t,
                                                                                  c_void_p, c_size_t,
                                                                                  c_size_t, c_float,
                                                                                  c_float]
        _lib.orc_spmm_bell_f64.argtypes = [c_size_t] * 5 + [c_void_p] * 3 + [c_size_t, c_void_p, c_size_t,
                                                                           c_double, c_double, c_void_p]
        _lib.orc_convert_from_f32.argtypes = [c_int, c_void_p, c_void_p, c_size_t]
        _lib.orc_convert_to_f32.argtypes = [c_int, c_void_p, c_void_p, c_size_t]
        _lib.orc_f32_to_f16.argtypes = [c_float]
        _lib.orc_f32_to_f16.restype = ctypes.c_uint16
        _lib.orc_f32_to_bf16.argtypes = [c_float]
        _lib.orc_f32_to_bf16.restype = ctypes.c_uint16
        _lib.orc_f16_to_f32.argtypes = [ctypes.c_uint16]
        _lib.orc_f16_to_f32.restype = c_float
        _lib.orc_bf16_to_f32.argtypes = [ctypes.c_uint16]
        _lib.orc_bf16_to_f32.restype = c_float
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def from_f32(dtype, x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.shape, dtype=np.uint16)
    lib().orc_convert_from_f32(dtype, _p(x), _p(out), x.size)
    return out


def to_f32(dtype, bits):
    bits = np.ascontiguousarray(bits, dtype=np.uint16)
    out = np.empty(bits.shape, dtype=np.float32)
    lib().orc_convert_to_f32(dtype, _p(bits), _p(out), bits.size)
    return out


def prune_blocks_ref(weights, m, n, blk_m=2, blk_n=2, sparsity_factor=0.5):
    """-> (pruned weights copy, uint64 mask).  `weights` any 2/4/8-byte dtype, m*n elements."""
    w = np.ascontiguousarray(weights).copy().reshape(-1)
    mask = np.empty(m * n, dtype=np.uint64)
    lib().orc_prune_blocks_ref(w.dtype.itemsize, _p(w), _p(mask), m, n, blk_m, blk_n, sparsity_factor)
    return w, mask


def prune24_strip(dtype, a_bits, want_mask=True):
    """a_bits: uint16 [rows, cols].  -> dict(dense, vals, meta, mask) in the CANONICAL layout."""
    a = np.ascontiguousarray(a_bits, dtype=np.uint16)
    rows, cols = a.shape
    G = (cols + 3) // 4
    dense = np.empty_like(a)
    vals = np.empty((rows, G * 2), dtype=np.uint16)
    meta = np.empty((rows, (G + 1) // 2), dtype=np.uint8)
    mask = np.empty(rows * cols, dtype=np.uint64) if want_mask else None
    lib().orc_prune24_strip(dtype, _p(a), cols, rows, cols, _p(dense), cols, _p(vals), _p(meta), _p(mask))
    return {"dense": dense, "vals": vals, "meta": meta,
            "mask": mask.reshape(rows, cols) if want_mask else None}


def prune24_tile(dtype, a_bits):
    a = np.ascontiguousarray(a_bits, dtype=np.uint16)
    rows, cols = a.shape
    dense = np.empty_like(a)
    mask = np.empty(rows * cols, dtype=np.uint64)
    lib().orc_prune24_tile(dtype, _p(a), cols, rows, cols, _p(dense), cols, _p(mask))
    return dense, mask.reshape(rows, cols)


def prune24_check(dtype, a_bits):
    a = np.ascontiguousarray(a_bits, dtype=np.uint16)
    return int(lib().orc_prune24_check(dtype, _p(a), a.shape[1], a.shape[0], a.shape[1]))


def pack_sm100(vals, meta, rows, cols):
    """CANONICAL -> SM100 device layout (uint8 buffers)."""
    mt, kt = (rows + 127) // 128, (cols + 127) // 128
    ov = np.zeros(mt * kt * 16384, dtype=np.uint8)
    om = np.zeros(mt * kt * 2048, dtype=np.uint8)
    vals = np.ascontiguousarray(vals)
    meta = np.ascontiguousarray(meta)
    lib().orc_pack_sm100(_p(vals), _p(meta), rows, cols, _p(ov), _p(om))
    return ov, om


def spmma_f64(dtype, a_dense_bits, b_bits, c_bits=None, alpha=1.0, beta=0.0, op_b=0):
    """fp64-accumulate D = alpha*A*op(B) + beta*C over storage-rounded inputs (A already pruned)."""
    a = np.ascontiguousarray(a_dense_bits, dtype=np.uint16)
    b = np.ascontiguousarray(b_bits, dtype=np.uint16)
    m, k = a.shape
    n = b.shape[0] if op_b else b.shape[1]
    c = np.ascontiguousarray(c_bits, dtype=np.uint16) if c_bits is not None else None
    out = np.empty((m, n), dtype=np.float64)
    lib().orc_spmma_f64(dtype, op_b, m, n, k, alpha, _p(a), k, _p(b), b.shape[1], beta, _p(c),
                        n if c is not None else 0, _p(out))
    return out


def spmma_compressed_f32(dtype, vals, meta, m, k, b_bits, alpha=1.0):
    """the timed CPU port: consumes the CANONICAL compressed operand, fp32 accumulate -> uint16 D"""
    b = np.ascontiguousarray(b_bits, dtype=np.uint16)
    n = b.shape[1]
    d = np.empty((m, n), dtype=np.uint16)
    lib().orc_spmma_compressed_f32(dtype, m, n, k, alpha, _p(vals), _p(meta), _p(b), n, 0.0, None, 0,
                                   _p(d), n)
    return d


def threshold_to_coo(dtype, a, threshold):
    a = np.ascontiguousarray(a)
    rows, cols = a.shape
    cap = rows * cols
    ri = np.empty(cap, dtype=np.int32)
    ci = np.empty(cap, dtype=np.int32)
    va = np.empty(cap, dtype=np.float32)
    rp = np.empty(rows + 1, dtype=np.int32)
    nnz = int(lib().orc_threshold_to_coo(dtype, _p(a), cols, rows, cols, threshold, _p(ri), _p(ci),
                                          _p(va), cap, _p(rp)))
    return ri[:nnz].copy(), ci[:nnz].copy(), va[:nnz].copy(), rp


def coo_to_csr(row_idx, rows):
    ri = np.ascontiguousarray(row_idx, dtype=np.int32)
    rp = np.empty(rows + 1, dtype=np.int32)
    lib().orc_coo_to_csr(_p(ri), ri.size, rows, _p(rp))
    return rp


def spmm_coo_batched_f64(m, k, n, nb, ri, ci, va, B, C=None, alpha=1.0, beta=0.0):
    """B: float32 [nb, n, k] (column-major k x n per batch); C likewise [nb, n, m]; -> f64 [nb, n, m]"""
    ri = np.ascontiguousarray(ri, dtype=np.int32)
    ci = np.ascontiguousarray(ci, dtype=np.int32)
    va = np.ascontiguousarray(va, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    Cc = np.ascontiguousarray(C, dtype=np.float32) if C is not None else None
    out = np.empty((nb, n, m), dtype=np.float64)
    lib().orc_spmm_coo_batched_f64(m, k, ri.size, n, nb, _p(ri), _p(ci), _p(va), _p(B), k, k * n, _p(Cc),
                                   m, m * n, alpha, beta, _p(out))
    return out


def spmm_coo_batched_f32(m, k, n, nb, ri, ci, va, B, C, alpha=1.0, beta=0.0):
    lib().orc_spmm_coo_batched_f32(m, k, ri.size, n, nb, _p(ri), _p(ci), _p(va), _p(B), k, k * n, _p(C), m,
                                   m * n, alpha, beta)
    return C


def spmm_bell_f64(rows, cols, n, block, ell_cols, col_idx, values, B, C=None, alpha=1.0, beta=0.0):
    col_idx = np.ascontiguousarray(col_idx, dtype=np.int64)
    values = np.ascontiguousarray(values, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    Cc = np.ascontiguousarray(C, dtype=np.float32) if C is not None else None
    out = np.empty((n, rows), dtype=np.float64)
    lib().orc_spmm_bell_f64(rows, cols, n, block, ell_cols, _p(col_idx), _p(values), _p(B), cols, _p(Cc),
                            rows, alpha, beta, _p(out))
    return out
