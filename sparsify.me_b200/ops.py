"""Host-side mirror of the reference's operator API on top of the C ABI (capi).

Names, argument meaning, defaults and return values follow the reference templates:
  sparsify            include/sparsify.me/sparsify.hxx:24-30  (weights, mask, m, n, sparsity_factor=0.5, stream)
  spmma               include/sparsify.me/spmma.hxx:21-33     -> [prune_ms, compress_ms, mul_ms] (:117)
  batched.spmm        include/sparsify.me/spmm.hxx:30-41      -> ms
  batched.strided_coo include/sparsify.me/spmm.hxx:140-153    -> ms
Tensors are torch CUDA tensors used as raw device buffers; all work is enqueued on torch's
current stream.  Nothing here computes on the host.
"""
import ctypes
import sys
from dataclasses import dataclass

import torch

from . import capi

_DT = {torch.float16: capi.F16, torch.bfloat16: capi.BF16, torch.float32: capi.F32,
       torch.float64: capi.F64}


def _dtype_code(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise capi.SpfyError(capi.E_UNSUPPORTED, "dtype", f"unsupported dtype {t.dtype}")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise capi.SpfyError(capi.E_INVALID, "device pointer",
                             "operand is not a CUDA tensor (sparsify.me_b200 has no host path)")
    return ctypes.c_void_p(t.data_ptr())


def _workspace(nbytes, device):
    """caller-provided scratch for one call (256-byte aligned like every torch allocation)"""
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


class _Timer:
    """cudaEvent pair on the current stream -- the reference's util::timer_t
    (include/sparsify.me/util/timer.hxx:24-55): begin/end record + synchronise."""

    def __init__(self):
        self.a = torch.cuda.Event(enable_timing=True)
        self.b = torch.cuda.Event(enable_timing=True)

    def begin(self):
        self.a.record()

    def end(self):
        self.b.record()
        self.b.synchronize()
        return self.a.elapsed_time(self.b)


# ----------------------------------------------------------------------------- A1
def sparsify(weights, mask, m, n, sparsity_factor=0.5, blk=(2, 2)):
    """sparsifyme::sparsify<BLK_M,BLK_N> (sparsify.hxx:24-82), exact positional semantics.
    `weights`: m*n elements (any 2/4/8-byte float type), pruned in place;
    `mask`: m*n int64/uint64 words (std::size_t in the reference)."""
    if mask.element_size() != 8:
        raise capi.SpfyError(capi.E_INVALID, "sparsify", "mask must hold 64-bit words")
    capi.spfy_prune_blocks_ref(_dtype_code(weights), _ptr(weights), _ptr(mask), m, n, blk[0], blk[1],
                               float(sparsity_factor), _stream())


# ------------------------------------------------------------------------- A2 / A3
@dataclass
class Compressed24:
    """The compressed 2:4 operand (what cusparseLtSpMMACompress returns, spmma.hxx:97-104)."""
    vals: torch.Tensor      # uint8 buffer
    meta: torch.Tensor      # uint8 buffer
    rows: int
    cols: int
    dtype: torch.dtype
    layout: int


def compressed_bytes(dtype, rows, cols, layout=capi.LAYOUT_SM100):
    vb, mb = ctypes.c_size_t(), ctypes.c_size_t()
    capi.spfy_compressed_bytes(_DT[dtype], rows, cols, layout, ctypes.byref(vb), ctypes.byref(mb))
    return vb.value, mb.value


def prune24(a, out_dense=None, mask=None, layout=capi.LAYOUT_SM100, mode=capi.PRUNE_STRIP_MAG,
            compress=True, inplace=False, out=None):
    """2:4 magnitude prune (+ compress) of the row-major matrix `a` [rows, cols] in ONE kernel.
    Replaces cusparseLtSpMMAPrune + cusparseLtSpMMACompress (spmma.hxx:85-104).
    Returns a Compressed24 (or None when compress=False)."""
    assert a.dim() == 2 and a.stride(1) == 1
    rows, cols = a.shape
    if inplace:
        out_dense = a
    comp = out
    if compress and comp is None:
        vb, mb = compressed_bytes(a.dtype, rows, cols, layout)
        comp = Compressed24(torch.empty(vb, dtype=torch.uint8, device=a.device),
                            torch.empty(mb, dtype=torch.uint8, device=a.device), rows, cols, a.dtype,
                            layout)
    capi.spfy_prune24(_dtype_code(a), mode, layout, _ptr(a), a.stride(0), _ptr(out_dense),
                      out_dense.stride(0) if out_dense is not None else 0,
                      _ptr(comp.vals) if comp else None, _ptr(comp.meta) if comp else None,
                      _ptr(mask), rows, cols, _stream())
    return comp


def prune24_check(a):
    """cusparseLtSpMMAPruneCheck (spmma.hxx:88-94): 0 iff `a` obeys 2:4 along its rows."""
    flag = torch.empty(1, dtype=torch.int32, device=a.device)
    capi.spfy_prune24_check(_dtype_code(a), _ptr(a), a.stride(0), a.shape[0], a.shape[1], _ptr(flag),
                            _stream())
    return int(flag.item())


def prune24_batched(mats, comps, layout=capi.LAYOUT_SM100, mode=capi.PRUNE_STRIP_MAG, inplace=False):
    """Prune+compress a whole list of weight matrices with as few launches as possible
    (spfy_prune24_batched).  `mats`: 2-D row-major tensors; `comps`: matching Compressed24.
    mode=PRUNE_TILE_MAG (what the reference asks cusparseLt for, spmma.hxx:86) prunes the matrices IN PLACE
    like cusparseLtSpMMAPrune(dA, dA) -- TILE needs the dense pruned matrix as an output; `inplace` asks for the
    same with STRIP."""
    items = (capi.Prune24Item * len(mats))()
    dense = inplace or mode == capi.PRUNE_TILE_MAG
    for i, (a, c) in enumerate(zip(mats, comps)):
        items[i] = capi.Prune24Item(a.data_ptr(), a.stride(0), a.data_ptr() if dense else None, a.stride(0) if dense else 0,
                                    c.vals.data_ptr(), c.meta.data_ptr(), a.shape[0], a.shape[1])
    capi.spfy_prune24_batched(_dtype_code(mats[0]), mode, layout, ctypes.cast(items, ctypes.c_void_p), len(mats),
                              _stream())


def alloc_compressed(dtype, rows, cols, device, layout=capi.LAYOUT_SM100):
    vb, mb = compressed_bytes(dtype, rows, cols, layout)
    return Compressed24(torch.empty(vb, dtype=torch.uint8, device=device),
                        torch.empty(mb, dtype=torch.uint8, device=device), rows, cols, dtype, layout)


def pack_compressed(comp):
    """Compressed24 -> one self-describing host buffer (bytes): the pruned-layer container of spfy_b200.h
    (header + values + metadata, checksummed).  The payload is the device image byte for byte."""
    vals = comp.vals.cpu().contiguous()
    meta = comp.meta.cpu().contiguous()
    n = ctypes.c_size_t()
    capi.spfy_packed_bytes(_DT[comp.dtype], comp.rows, comp.cols, comp.layout, ctypes.byref(n))
    buf = ctypes.create_string_buffer(n.value)
    capi.spfy_packed_write(_DT[comp.dtype], comp.layout, comp.rows, comp.cols, vals.data_ptr(), meta.data_ptr(),
                           ctypes.cast(buf, ctypes.c_void_p), n.value)
    return buf.raw


def unpack_compressed(blob, device):
    """bytes written by pack_compressed -> Compressed24 on `device` (validated: magic, version, sizes, checksum)"""
    buf = ctypes.create_string_buffer(blob, len(blob))
    dt, layout = ctypes.c_int(), ctypes.c_int()
    rows, cols, vo, vb, mo, mb = (ctypes.c_size_t() for _ in range(6))
    capi.spfy_packed_read(ctypes.cast(buf, ctypes.c_void_p), len(blob), ctypes.byref(dt), ctypes.byref(layout),
                          ctypes.byref(rows), ctypes.byref(cols), ctypes.byref(vo), ctypes.byref(vb), ctypes.byref(mo),
                          ctypes.byref(mb))
    raw = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
    tdt = {v: k for k, v in _DT.items()}[dt.value]
    return Compressed24(raw[vo.value: vo.value + vb.value].to(device), raw[mo.value: mo.value + mb.value].to(device),
                        rows.value, cols.value, tdt, layout.value)


class SpmmaPlan:
    """spfy_spmma_plan_*: a list of independent D_i = alpha_i * A_i(2:4) * op(B_i) + beta_i * C_i executed
    by one persistent launch per ring-geometry class present (tensor maps + tile schedule built once, like
    cusparseLtMatmulPlanInit at spmma.hxx:79).  `problems`: dicts with comp, b, out and optional c, alpha,
    beta, op_b, out_t (out is then [n, m], the transposed result; beta must be 0).  The plan keeps the tensors alive.
    Optional `replicas` per problem: device addresses (ints) of further copies of `out` (same shape and row pitch) that
    the epilogue stores every tile to as well -- e.g. the slab of this rank in the gather arena of every peer GPU
    (spfy_spmma_plan_create_replicated: the fused output gather); every problem must name the same number of them.
    A problem with `conv=(kh, kw, stride, pad)` is a convolution layer (spfy_spmma_plan_create_conv): `b` is then the
    NHWC activation tensor [batch, h, w, ch] and comp holds the weights with K in (kh, kw, ch) order; the unfolded
    operand is never built.  `out` is [m, batch*ho*wo], or NHWC [batch, ho, wo, m] with out_t."""

    def __init__(self, problems):
        self._keep = problems
        n = len(problems)
        arr = (capi.SpmmaProblem * n)()
        convs = (capi.ConvDesc * max(n, 1))()
        conv_ptrs = (ctypes.c_void_p * max(n, 1))()
        any_conv = False
        dtype = None
        for i, q in enumerate(problems):
            comp, b, out = q["comp"], q["b"], q["out"]
            c = q.get("c")
            op_b = q.get("op_b", capi.OP_N)
            dtype = comp.dtype
            ldb = b.stride(0)
            ldd = out.stride(0)
            if q.get("conv"):
                kh, kw, stride, pad = q["conv"]
                assert b.dim() == 4 and b.is_contiguous()
                nb, h, w, ch = b.shape
                assert comp.cols == kh * kw * ch
                convs[i] = capi.ConvDesc(nb, h, w, ch, kh, kw, stride, pad)
                conv_ptrs[i] = ctypes.addressof(convs[i])
                any_conv = True
                op_b, ldb = capi.OP_T, comp.cols
                nn = nb * ((h + 2 * pad - kh) // stride + 1) * ((w + 2 * pad - kw) // stride + 1)
                if q.get("out_t"):
                    assert out.is_contiguous()
                    ldd = comp.rows
            else:
                nn = b.shape[1] if op_b == capi.OP_N else b.shape[0]
            if q.get("out_t"):  # `out` is [n, m]: the transposed result
                op_b |= capi.OUT_T
            arr[i] = capi.SpmmaProblem(op_b, comp.rows, nn, comp.cols, comp.vals.data_ptr(), comp.meta.data_ptr(),
                                       b.data_ptr(), ldb, c.data_ptr() if c is not None else None,
                                       c.stride(0) if c is not None else 0, out.data_ptr(), ldd,
                                       float(q.get("alpha", 1.0)), float(q.get("beta", 0.0)))
        self._h = ctypes.c_void_p()
        nrep = len(problems[0].get("replicas", ())) if n else 0
        if any(len(q.get("replicas", ())) != nrep for q in problems):
            raise ValueError("SpmmaPlan: every problem needs the same number of replicas")
        if any_conv:
            if nrep:
                raise ValueError("SpmmaPlan: replicas and convolution problems cannot be combined")
            capi.spfy_spmma_plan_create_conv(_DT[dtype], ctypes.cast(arr, ctypes.c_void_p), ctypes.cast(conv_ptrs, ctypes.c_void_p),
                                             n, ctypes.byref(self._h))
        elif nrep:
            rep = (ctypes.c_void_p * (n * nrep))(*[int(a) for q in problems for a in q["replicas"]])
            capi.spfy_spmma_plan_create_replicated(_DT[dtype], ctypes.cast(arr, ctypes.c_void_p), n, nrep,
                                                   ctypes.cast(rep, ctypes.c_void_p), ctypes.byref(self._h))
        else:
            capi.spfy_spmma_plan_create(_DT[dtype], ctypes.cast(arr, ctypes.c_void_p), n, ctypes.byref(self._h))
        self.replicas = nrep
        self.launches = capi.spfy_spmma_plan_launches(self._h)

    def run(self):
        capi.spfy_spmma_plan_run(self._h, _stream())

    def run_launch(self, index):
        capi.spfy_spmma_plan_run_launch(self._h, index, _stream())

    def launch_info(self, index):
        v = [ctypes.c_int() for _ in range(4)]
        capi.spfy_spmma_plan_launch_info(self._h, index, *[ctypes.byref(x) for x in v])
        return dict(zip(("problems", "units", "stages", "smem_bytes"), (x.value for x in v)))

    def close(self):
        if self._h:
            capi.spfy_spmma_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


# ----------------------------------------------------------------------------- A4
def spmma_compressed(comp, b, c=None, out=None, alpha=1.0, beta=0.0, op_b=capi.OP_N, out_t=False):
    """D = alpha * A(2:4) * op(B) + beta * C on tcgen05.mma.sp (cusparseLtMatmul, spmma.hxx:106-114).
    All dense operands row-major.  Returns D (== out, or a new tensor); out_t: D is [n, m], the transposed result
    (SPFY_OUT_T: NHWC for a convolution layer; beta must be 0)."""
    assert comp.layout == capi.LAYOUT_SM100
    m, k = comp.rows, comp.cols
    n = b.shape[1] if op_b == capi.OP_N else b.shape[0]
    assert (b.shape[0] if op_b == capi.OP_N else b.shape[1]) == k
    if out is None:
        out = torch.empty((n, m) if out_t else (m, n), dtype=comp.dtype, device=b.device)
    if out_t:
        op_b |= capi.OUT_T
    capi.spfy_spmma(_DT[comp.dtype], op_b, m, n, k, float(alpha), _ptr(comp.vals), _ptr(comp.meta),
                    _ptr(b), b.stride(0), float(beta), _ptr(c), c.stride(0) if c is not None else 0,
                    _ptr(out), out.stride(0), None, 0, _stream())
    return out


def permute_conv_weights(w, c, kh, kw):
    """[m, c*kh*kw] conv weights in (c, kh, kw) column order (a flattened torch conv weight, the K order of `unfold`)
    -> [m, kh*kw*c] in (kh, kw, c) order, the K order of spmma_conv.  Do this BEFORE pruning."""
    out = torch.empty_like(w)
    capi.spfy_permute_conv_weights(_ptr(w), _ptr(out), w.shape[0], c, kh, kw, _stream())
    return out


def spmma_conv(comp, x, kh, kw, stride=1, pad=0, c=None, out=None, alpha=1.0, beta=0.0):
    """Implicit GEMM (spfy_spmma_conv): D[m, batch*ho*wo] = alpha * A(2:4) * im2col(x) + beta * C with x the NHWC
    activation tensor [batch, h, w, ch] -- the unfolded K x N operand of datasets/get_shapes.py:29-41 is never built.
    `comp`: compressed weights whose K is ordered (kh, kw, ch) (permute_conv_weights before prune24)."""
    assert x.dim() == 4 and x.is_contiguous()
    nb, h, w, ch = x.shape
    ho, wo = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
    assert comp.cols == kh * kw * ch
    n = nb * ho * wo
    if out is None:
        out = torch.empty(comp.rows, n, dtype=comp.dtype, device=x.device)
    desc = capi.ConvDesc(nb, h, w, ch, kh, kw, stride, pad)
    capi.spfy_spmma_conv(_DT[comp.dtype], ctypes.byref(desc), comp.rows, float(alpha), _ptr(comp.vals), _ptr(comp.meta),
                         _ptr(x), float(beta), _ptr(c), c.stride(0) if c is not None else 0, _ptr(out), out.stride(0),
                         _stream())
    return out


def spmma_conv_nhwc(comp, x, kh, kw, stride=1, pad=0, out=None, alpha=1.0):
    """spfy_spmma_conv_nhwc: the same implicit GEMM with the result in NHWC, [batch, ho, wo, m] -- the `x` of the next
    layer, so a chain of convolutions never transposes or unfolds anything."""
    assert x.dim() == 4 and x.is_contiguous()
    nb, h, w, ch = x.shape
    ho, wo = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
    assert comp.cols == kh * kw * ch
    if out is None:
        out = torch.empty(nb, ho, wo, comp.rows, dtype=comp.dtype, device=x.device)
    assert out.is_contiguous() and out.numel() == nb * ho * wo * comp.rows
    desc = capi.ConvDesc(nb, h, w, ch, kh, kw, stride, pad)
    capi.spfy_spmma_conv_nhwc(_DT[comp.dtype], ctypes.byref(desc), comp.rows, float(alpha), _ptr(comp.vals),
                              _ptr(comp.meta), _ptr(x), _ptr(out), comp.rows, _stream())
    return out


def spmma(a, b, c, m, n, k, batch_size=1, transpose_a=capi.OP_N, transpose_b=capi.OP_N, alpha=1.0,
          beta=0.0, prune_mode=capi.PRUNE_TILE_MAG):
    """sparsifyme::spmma (spmma.hxx:21-118): prune A in place (2:4 magnitude), compress, multiply
    into C (D aliases C, :52-53).  Returns [prune_ms, compress_ms, mul_ms] like :117.  Like the header
    (include/sparsify.me/spmma.hxx) it prunes with TILE_MAG, the algorithm the reference requests from
    cusparseLt (:86); prune_mode=PRUNE_STRIP_MAG is the header's -DSPARSIFYME_PRUNE_STRIP.  Prune and
    compress are timed together under `prune`; `compress` only times the (empty) remainder;
    batch_size is accepted and unused exactly like the reference (:29)."""
    del batch_size
    if transpose_a != capi.OP_N:
        raise capi.SpfyError(capi.E_UNSUPPORTED, "spmma", "transpose_a is not supported")
    if m % 8 or n % 8 or k % 8:  # spmma.hxx:45-49 (warning only, execution continues)
        print("Invalid matrix sizes for data type __half. Rows and columns must be divisible by 8.", file=sys.stderr)
    a2 = a.view(-1)[: m * k].view(m, k)
    b2 = b.view(-1)[: k * n].view(k, n) if transpose_b == capi.OP_N else b.view(-1)[: k * n].view(n, k)
    c2 = c.view(-1)[: m * n].view(m, n)
    vb, mb = compressed_bytes(a.dtype, m, k)
    t = _Timer()
    t.begin()
    comp = Compressed24(torch.empty(vb, dtype=torch.uint8, device=a.device),
                        torch.empty(mb, dtype=torch.uint8, device=a.device), m, k, a.dtype,
                        capi.LAYOUT_SM100)
    prune24(a2, inplace=True, out=comp, mode=prune_mode)
    prune_ms = t.end()
    t.begin()
    compress_ms = t.end()
    t.begin()
    spmma_compressed(comp, b2, c=c2 if beta != 0.0 else None, out=c2, alpha=alpha, beta=beta,
                     op_b=transpose_b)
    mul_ms = t.end()
    return [prune_ms, compress_ms, mul_ms]


# ------------------------------------------------------------------- unstructured path
def threshold_to_coo(a, threshold, capacity=None, want_csr=False, sync=True):
    """Unstructured magnitude prune: keep x iff |x| > threshold; COO sorted by (row, col),
    fp32 values, int32 indices (the operand format of spmm.hxx:165-168).
    Returns (row_idx, col_idx, vals, nnz[, row_ptr]); index/value tensors are trimmed to nnz.
    sync=False skips the nnz read-back: the arrays keep their full capacity, nnz is a device tensor and
    row_ptr is always returned (feed it to batched.csr, which needs no host-side nnz)."""
    rows, cols = a.shape
    cap = rows * cols if capacity is None else capacity
    dev = a.device
    ri = torch.empty(cap, dtype=torch.int32, device=dev)
    ci = torch.empty(cap, dtype=torch.int32, device=dev)
    va = torch.empty(cap, dtype=torch.float32, device=dev)
    nnz = torch.zeros(1, dtype=torch.int64, device=dev)
    rp = torch.empty(rows + 1, dtype=torch.int32, device=dev)
    wb = ctypes.c_size_t()
    capi.spfy_threshold_workspace_bytes(rows, cols, ctypes.byref(wb))
    ws = torch.empty(max(wb.value, 16), dtype=torch.uint8, device=dev)
    capi.spfy_threshold_to_coo(_dtype_code(a), _ptr(a), a.stride(0), rows, cols, float(threshold),
                               _ptr(ri), _ptr(ci), _ptr(va), cap, _ptr(nnz), _ptr(rp), _ptr(ws),
                               ws.numel(), _stream())
    if not sync:
        return ri, ci, va, nnz, rp
    n = int(nnz.item())
    kept = min(n, cap)
    res = (ri[:kept], ci[:kept], va[:kept], n)
    return res + (rp,) if want_csr else res


def coo_to_csr(row_idx, rows):
    rp = torch.empty(rows + 1, dtype=torch.int32, device=row_idx.device)
    capi.spfy_coo_to_csr(_ptr(row_idx), row_idx.numel(), rows, _ptr(rp), _stream())
    return rp


class batched:
    """namespace sparsifyme::batched"""

    @staticmethod
    def strided_coo(a_num_rows, a_num_cols, a_nnz, b_num_rows, b_num_cols, num_batches, a_rows, a_cols,
                    a_values, b, c, alpha=1.0, beta=0.0, alg=capi.SPMM_ALG_DEFAULT):
        """batched::strided_coo (spmm.hxx:140-193): C_b = alpha*A*B_b + beta*C_b with ONE COO A
        (row-sorted), B_b = b[i] k x n column-major (ldb = k), C_b m x n column-major (ldc = m);
        b and c are single slabs strided by ldb*n / ldc*n (:172,:175 intent).  Returns ms.
        `alg`: SPMM_ALG_* of spfy_b200.h (DEFAULT: tensor cores when A is dense enough to pay)."""
        assert b_num_rows == a_num_cols
        ldb, ldc = b_num_rows, a_num_rows
        wb = ctypes.c_size_t()
        capi.spfy_spmm_workspace_bytes(alg, a_num_rows, a_num_cols, b_num_cols, num_batches, a_nnz, ctypes.byref(wb))
        ws = _workspace(wb.value, b.device)
        t = _Timer()
        t.begin()
        capi.spfy_spmm_coo_strided_batched(alg, a_num_rows, a_num_cols, a_nnz, b_num_cols, num_batches,
                                           _ptr(a_rows), _ptr(a_cols), _ptr(a_values), _ptr(b), ldb,
                                           ldb * b_num_cols, _ptr(c), ldc, ldc * b_num_cols,
                                           float(alpha), float(beta), _ptr(ws), ws.numel(), _stream())
        return t.end()

    @staticmethod
    def csr(m, k, n, num_batches, row_ptr, col_idx, vals, b, c, alpha=1.0, beta=0.0, alg=capi.SPMM_ALG_DEFAULT):
        wb = ctypes.c_size_t()
        capi.spfy_spmm_workspace_bytes(alg, m, k, n, num_batches, col_idx.numel(), ctypes.byref(wb))
        ws = _workspace(wb.value, b.device)
        capi.spfy_spmm_csr_strided_batched(alg, m, k, n, num_batches, _ptr(row_ptr), _ptr(col_idx),
                                           _ptr(vals), _ptr(b), k, k * n, _ptr(c), m, m * n,
                                           float(alpha), float(beta), _ptr(ws), ws.numel(), _stream())

    @staticmethod
    def spmm(col_idx_list, values_list, b, c_list, m, n, k, block, ell_cols, alpha=1.0, beta=0.0,
             alg=capi.SPMM_ALG_DEFAULT):
        """batched::spmm (spmm.hxx:30-138): per batch C_b = alpha*A_b*B + beta*C_b, A_b blocked-ELL
        (containers/ell.hxx:24-33), B k x n column-major shared, C_b m x n column-major.
        One launch covers every batch element (the reference fans out one host thread +
        stream per batch, :94-115).  Returns ms."""
        dev = b.device
        nb = len(values_list)
        ci = torch.tensor([t.data_ptr() for t in col_idx_list], dtype=torch.int64, device=dev)
        va = torch.tensor([t.data_ptr() for t in values_list], dtype=torch.int64, device=dev)
        cs = torch.tensor([t.data_ptr() for t in c_list], dtype=torch.int64, device=dev)
        wb = ctypes.c_size_t()
        capi.spfy_spmm_bell_workspace_bytes(alg, _dtype_code(b), m, k, n, nb, ctypes.byref(wb))
        ws = _workspace(wb.value, dev)
        t = _Timer()
        t.begin()
        capi.spfy_spmm_bell_batched(alg, _dtype_code(b), m, k, n, block, ell_cols, nb, _ptr(ci), _ptr(va),
                                    _ptr(b), k, _ptr(cs), m, float(alpha), float(beta), _ptr(ws), ws.numel(),
                                    _stream())
        return t.end()


    @staticmethod
    def gemm(a, b, c, m, n, k, transpose_a=capi.OP_N, transpose_b=capi.OP_N, alpha=1.0, beta=0.0,
             precision=capi.GEMM_PRECISE, strided=True):
        """batched::gemm (gemm.hxx:25-195): column-major C_i[m x n] = alpha*op(A_i)*op(B_i) + beta*C_i on tcgen05
        (fp16 / bf16: kind::f16; fp32: 3xTF32, or one TF32 product with precision=GEMM_FAST).
        a: [nb, ...] slab of column-major operands (or a single matrix shared by the batch), same for b;
        c: [nb, n, m] slab (column-major m x n per batch).  lda = m (k if transposed), ldb = k (n), ldc = m (:79-81).
        strided=False goes through the pointer-array entry (what the header template calls).  Returns ms."""
        nb = c.shape[0]
        lda = m if transpose_a == capi.OP_N else k
        ldb = k if transpose_b == capi.OP_N else n
        sa = a.stride(0) if a.dim() == 3 and a.shape[0] == nb and nb > 1 else 0
        sb = b.stride(0) if b.dim() == 3 and b.shape[0] == nb and nb > 1 else 0
        t = _Timer()
        wb = ctypes.c_size_t()
        capi.spfy_gemm_workspace_bytes(_dtype_code(c), transpose_a, transpose_b, m, n, k, lda, ldb, nb, ctypes.byref(wb))
        ws = _workspace(wb.value, c.device)
        if strided:
            t.begin()
            capi.spfy_gemm_strided_batched(_dtype_code(c), precision, transpose_a, transpose_b, m, n, k, float(alpha),
                                           _ptr(a), lda, sa, _ptr(b), ldb, sb, float(beta), _ptr(c), m, c.stride(0), nb,
                                           _ptr(ws), ws.numel(), _stream())
            return t.end()
        es = a.element_size()
        pa = (ctypes.c_void_p * nb)(*[a.data_ptr() + i * sa * es for i in range(nb)])
        pb = (ctypes.c_void_p * nb)(*[b.data_ptr() + i * sb * es for i in range(nb)])
        pc = (ctypes.c_void_p * nb)(*[c.data_ptr() + i * c.stride(0) * es for i in range(nb)])
        t.begin()
        capi.spfy_gemm_batched(_dtype_code(c), precision, transpose_a, transpose_b, m, n, k, float(alpha),
                               ctypes.cast(pa, ctypes.c_void_p), lda, ctypes.cast(pb, ctypes.c_void_p), ldb, float(beta),
                               ctypes.cast(pc, ctypes.c_void_p), m, nb, _ptr(ws), ws.numel(), _stream())
        return t.end()
