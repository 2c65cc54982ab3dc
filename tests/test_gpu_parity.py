"""GPU parity tests: every CUDA entry point (through the C ABI) against the CPU oracle on the same
seeded inputs.  Bit-exact for masks / compressed values / metadata / indices; GEMM outputs within
max relative error 1e-2 of the fp64 oracle (north_star tolerance for fp16/bf16)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL_TOL = 1e-2  # BASELINE.json north_star: "max relative error 1e-2 for fp16/bf16"


def rand_bits(orc, dtype, shape, seed, lo=-1.0, hi=1.0):
    rng = np.random.default_rng(seed)
    return orc.from_f32(dtype, rng.uniform(lo, hi, shape).astype(np.float32))


def to_dev(bits, dtype_code, dev):
    tdt = torch.float16 if dtype_code == 0 else torch.bfloat16
    return torch.from_numpy(np.ascontiguousarray(bits).view(np.int16).copy()).to(dev).view(tdt)


def bits_of(t):
    return t.contiguous().view(torch.int16).cpu().numpy().view(np.uint16)


def rel_err(got, want):
    scale = np.maximum(np.abs(want), 1e-2 * max(np.abs(want).max(), 1e-30))
    return float(np.max(np.abs(got - want) / scale))


# ------------------------------------------------------------------ A1 positional sparsify
@pytest.mark.parametrize("m,n", [(8, 8), (12544, 147), (196, 4608), (7, 5), (1, 1), (64, 64)])
@pytest.mark.parametrize("np_dtype,t_dtype", [(np.float32, torch.float32), (np.float16, torch.float16),
                                              (np.float64, torch.float64)])
def test_sparsify_positional_matches_oracle(spfy, orc, cuda, m, n, np_dtype, t_dtype):
    w = (1.0 + (np.arange(m * n) % 251)).astype(np_dtype)
    want_w, want_mask = orc.prune_blocks_ref(w, m, n, 2, 2, 0.5)
    dw = torch.from_numpy(w.copy()).to(cuda)
    dmask = torch.full((m * n,), 77, dtype=torch.int64, device=cuda)
    spfy.sparsify(dw, dmask, m, n)
    assert np.array_equal(dw.cpu().numpy(), want_w)
    assert np.array_equal(dmask.cpu().numpy().view(np.uint64), want_mask)


@pytest.mark.parametrize("blk,sf", [((2, 2), 0.25), ((2, 2), 0.75), ((4, 4), 0.5), ((4, 2), 0.5),
                                    ((2, 4), 0.5), ((1, 1), 0.5), ((1, 1), 1.0), ((2, 2), 1.0)])
def test_sparsify_other_blocks(spfy, orc, cuda, blk, sf):
    m, n = 36, 52
    w = (1.0 + (np.arange(m * n) % 251)).astype(np.float32)
    want_w, want_mask = orc.prune_blocks_ref(w, m, n, blk[0], blk[1], sf)
    dw = torch.from_numpy(w.copy()).to(cuda)
    dmask = torch.zeros(m * n, dtype=torch.int64, device=cuda)
    spfy.sparsify(dw, dmask, m, n, sparsity_factor=sf, blk=blk)
    assert np.array_equal(dw.cpu().numpy(), want_w)
    assert np.array_equal(dmask.cpu().numpy().view(np.uint64), want_mask)


# ------------------------------------------------------------------ A2/A3 prune + compress
PRUNE_SHAPES = [(64, 147), (128, 128), (64, 64), (256, 2304), (512, 4608), (130, 260), (1, 4), (3, 1),
                (2048, 512), (100, 1000), (5, 37)]


@pytest.mark.parametrize("rows,cols", PRUNE_SHAPES)
@pytest.mark.parametrize("dt", [0, 1])
def test_prune24_canonical_bit_exact(spfy, orc, cuda, rows, cols, dt):
    bits = rand_bits(orc, dt, (rows, cols), seed=rows * 7919 + cols)
    want = orc.prune24_strip(dt, bits)
    a = to_dev(bits, dt, cuda)
    dense = torch.empty_like(a)
    mask = torch.zeros(rows * cols, dtype=torch.int64, device=cuda)
    comp = spfy.prune24(a, out_dense=dense, mask=mask, layout=spfy.LAYOUT_CANONICAL)
    assert np.array_equal(bits_of(dense), want["dense"])
    assert np.array_equal(mask.cpu().numpy().view(np.uint64).reshape(rows, cols), want["mask"])
    assert np.array_equal(comp.vals.cpu().numpy().view(np.uint16).reshape(rows, -1), want["vals"])
    assert np.array_equal(comp.meta.cpu().numpy().reshape(rows, -1), want["meta"])
    assert spfy.prune24_check(dense) == 0
    assert orc.prune24_check(dt, want["dense"]) == 0


@pytest.mark.parametrize("rows,cols", PRUNE_SHAPES)
def test_prune24_sm100_layout_bit_exact(spfy, orc, cuda, rows, cols):
    bits = rand_bits(orc, 0, (rows, cols), seed=rows * 31 + cols)
    want = orc.prune24_strip(0, bits)
    ov, om = orc.pack_sm100(want["vals"], want["meta"], rows, cols)
    a = to_dev(bits, 0, cuda)
    comp = spfy.prune24(a, layout=spfy.LAYOUT_SM100)
    assert np.array_equal(comp.vals.cpu().numpy(), ov)
    assert np.array_equal(comp.meta.cpu().numpy(), om)


@pytest.mark.parametrize("rows,cols", [(128, 128), (256, 2304), (200, 96), (130, 272), (2048, 512), (1, 16), (129, 4608)])
@pytest.mark.parametrize("dense", [False, True])
@pytest.mark.parametrize("dt", [0, 1])
def test_prune24_fast_path_both_layouts(spfy, orc, cuda, rows, cols, dense, dt):
    """cols % 16 == 0 and no mask: the packed-select kernel (prune24_fast_kernel), with and without the
    pruned dense copy, tie-rich inputs included."""
    bits = rand_bits(orc, dt, (rows, cols), seed=rows * 131 + cols + dense)
    bits[::3, : cols // 2] &= 0xFC00  # few distinct magnitudes: plenty of ties inside groups
    want = orc.prune24_strip(dt, bits)
    ov, om = orc.pack_sm100(want["vals"], want["meta"], rows, cols)
    a = to_dev(bits, dt, cuda)
    for layout in (spfy.LAYOUT_SM100, spfy.LAYOUT_CANONICAL):
        out = torch.empty_like(a) if dense else None
        comp = spfy.prune24(a, out_dense=out, layout=layout)
        if dense:
            assert np.array_equal(bits_of(out), want["dense"])
        if layout == spfy.LAYOUT_SM100:
            assert np.array_equal(comp.vals.cpu().numpy(), ov)
            assert np.array_equal(comp.meta.cpu().numpy(), om)
        else:
            assert np.array_equal(comp.vals.cpu().numpy().view(np.uint16).reshape(rows, -1), want["vals"])
            assert np.array_equal(comp.meta.cpu().numpy().reshape(rows, -1), want["meta"])
    # in place (the reference's spmma mutates A: spmma.hxx:86)
    spfy.prune24(a, inplace=True, compress=False)
    assert np.array_equal(bits_of(a), want["dense"])


@pytest.mark.parametrize("rows,cols", [(64, 256), (33, 148), (128, 4608)])
def test_prune24_check_flags_violations(spfy, orc, cuda, rows, cols):
    """cusparseLtSpMMAPruneCheck semantics (spmma.hxx:88-94): 0 for a valid 2:4 matrix, non-zero as soon as
    one group of four keeps three values; -0 counts as zero; both the 128-bit and the scalar path."""
    bits = rand_bits(orc, 0, (rows, cols), seed=rows + cols)
    pruned = orc.prune24_strip(0, bits, want_mask=False)["dense"]
    d = to_dev(pruned, 0, cuda)
    assert spfy.prune24_check(d) == 0 and orc.prune24_check(0, pruned) == 0
    bad = pruned.copy()
    r, g = rows - 1, (cols // 4) - 1
    bad[r, 4 * g: 4 * g + 3] = 0x3C00  # three ones in the last full group of the last row
    assert orc.prune24_check(0, bad) != 0
    assert spfy.prune24_check(to_dev(bad, 0, cuda)) != 0
    negz = pruned.copy()
    negz[0, :4] = [0x8000, 0x3C00, 0x8000, 0x3C00]  # two values + two negative zeros: still valid
    assert spfy.prune24_check(to_dev(negz, 0, cuda)) == 0 and orc.prune24_check(0, negz) == 0


def test_prune24_inplace_and_strided(spfy, orc, cuda):
    rows, cols, ld = 96, 200, 256
    bits = rand_bits(orc, 0, (rows, ld), seed=5)
    want = orc.prune24_strip(0, bits[:, :cols])
    buf = to_dev(bits, 0, cuda)
    a = buf[:, :cols]
    spfy.prune24(a, inplace=True, compress=False)
    got = bits_of(buf)
    assert np.array_equal(got[:, :cols], want["dense"])
    assert np.array_equal(got[:, cols:], bits[:, cols:])  # padding columns untouched


def test_prune24_special_values(spfy, orc, cuda):
    """ties, +-0, inf, nan, denormals: the tie-break (lower index wins) and the key order
    (NaN > Inf > finite, -0 == +0) must agree bit for bit."""
    specials = np.array([0x0000, 0x8000, 0x7C00, 0xFC00, 0x7E00, 0xFE01, 0x0001, 0x8001, 0x3C00, 0xBC00,
                         0x3C01, 0x7BFF, 0xFBFF, 0x0400, 0x83FF, 0x3800], dtype=np.uint16)
    rng = np.random.default_rng(11)
    bits = rng.choice(specials, size=(64, 256)).astype(np.uint16)
    for dt in (0, 1):
        want = orc.prune24_strip(dt, bits)
        a = to_dev(bits, dt, cuda)
        dense = torch.empty_like(a)
        comp = spfy.prune24(a, out_dense=dense, layout=spfy.LAYOUT_CANONICAL)
        assert np.array_equal(bits_of(dense), want["dense"])
        assert np.array_equal(comp.meta.cpu().numpy().reshape(64, -1), want["meta"])
        assert np.array_equal(comp.vals.cpu().numpy().view(np.uint16).reshape(64, -1), want["vals"])


def test_prune24_all_ties_keeps_first_two(spfy, cuda):
    a = torch.ones(4, 16, dtype=torch.float16, device=cuda)
    dense = torch.empty_like(a)
    spfy.prune24(a, out_dense=dense, compress=False)
    want = torch.tensor([1, 1, 0, 0] * 4, dtype=torch.float16, device=cuda).expand(4, 16)
    assert torch.equal(dense, want)


def test_prune24_idempotent_full_size(spfy, cuda):
    """size-independent property at a BASELINE-size operand: pruning a pruned matrix is a no-op,
    every group of 4 keeps exactly 2 slots, and kept entries are unchanged."""
    torch.manual_seed(3)
    a = (torch.rand(512, 4608, device=cuda) * 2 - 1).half()
    d1 = torch.empty_like(a)
    c1 = spfy.prune24(a, out_dense=d1)
    d2 = torch.empty_like(a)
    c2 = spfy.prune24(d1, out_dense=d2)
    assert torch.equal(d1, d2)
    assert torch.equal(c1.vals, c2.vals) and torch.equal(c1.meta, c2.meta)
    nz = (d1 != 0).view(512, -1, 4).sum(-1)
    assert int(nz.max()) <= 2
    kept = d1 != 0
    assert torch.equal(d1[kept], a[kept])
    # magnitude property: every dropped entry is <= the smaller kept entry of its group
    g = a.abs().float().view(512, -1, 4)
    top2 = g.topk(2, dim=-1).values[..., 1]
    dropped = torch.where(kept.view(512, -1, 4), torch.zeros_like(g), g)
    assert bool((dropped.max(-1).values <= top2).all())


def test_prune24_tile_mode_matches_oracle(spfy, orc, cuda):
    bits = rand_bits(orc, 0, (64, 96), seed=21)
    want_dense, _ = orc.prune24_tile(0, bits)
    a = to_dev(bits, 0, cuda)
    dense = torch.empty_like(a)
    spfy.prune24(a, out_dense=dense, mode=spfy.PRUNE_TILE_MAG, compress=False)
    got = bits_of(dense)
    assert np.array_equal(got, want_dense)
    nz = (got.reshape(16, 4, 24, 4) != 0)
    assert (nz.sum(axis=3) <= 2).all() and (nz.sum(axis=1) <= 2).all()


def _tile_fixtures():
    import glob
    import os
    return sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tile_*.npz")))


@pytest.mark.parametrize("path", _tile_fixtures(), ids=lambda p: p.rsplit("/", 1)[-1])
def test_prune24_tile_kernel_is_bit_exact_with_cusparselt(spfy, cuda, path):
    """the CUDA TILE kernel against what cusparseLt 0.7.1 (the library behind spmma.hxx:86) chose on the probe
    tiles: every exact tie set, tie-rich integers and fp32-rounding cases -- no oracle in between"""
    from test_golden import tile_fixture, tiles_as_matrix
    dt, tiles, pattern = tile_fixture(path)
    code = 0 if dt == "f16" else 1
    keep = ((pattern[:, None] >> np.arange(16)) & 1).astype(bool)
    want = tiles_as_matrix(np.where(keep, tiles, 0).astype(np.uint16))
    a = to_dev(tiles_as_matrix(tiles), code, cuda)
    dense = torch.empty_like(a)
    spfy.prune24(a, out_dense=dense, mode=spfy.PRUNE_TILE_MAG, compress=False)
    assert np.array_equal(bits_of(dense), want)
    # in place (the reference prunes dA into dA: spmma.hxx:86)
    spfy.prune24(a, out_dense=a, mode=spfy.PRUNE_TILE_MAG, compress=False)
    assert np.array_equal(bits_of(a), want)


@pytest.mark.parametrize("dtype", [0, 1])
@pytest.mark.parametrize("shape", [(64, 96), (130, 262), (7, 5), (4, 4), (1024, 512)])
def test_prune24_tile_ragged_and_compress(spfy, orc, cuda, dtype, shape):
    """ragged edges are padded with +0 like the oracle; TILE + compress equals STRIP-compress of the TILE result"""
    bits = rand_bits(orc, dtype, shape, seed=77 + shape[0])
    want_dense, _ = orc.prune24_tile(dtype, bits)
    a = to_dev(bits, dtype, cuda)
    dense = torch.empty_like(a)
    comp = spfy.prune24(a, out_dense=dense, mode=spfy.PRUNE_TILE_MAG, layout=spfy.LAYOUT_CANONICAL)
    assert np.array_equal(bits_of(dense), want_dense)
    ref = orc.prune24_strip(dtype, want_dense, want_mask=False)
    assert np.array_equal(comp.vals.cpu().numpy().view(np.uint16).reshape(shape[0], -1), ref["vals"])
    assert np.array_equal(comp.meta.cpu().numpy().reshape(shape[0], -1), ref["meta"])
    # column-strided views take the scalar path (rows not 8-byte aligned)
    wide = torch.zeros(shape[0], shape[1] + 3, dtype=a.dtype, device=cuda)
    wide[:, 1:1 + shape[1]] = a
    view = wide[:, 1:1 + shape[1]]
    spfy.prune24(view, out_dense=view, mode=spfy.PRUNE_TILE_MAG, compress=False)
    assert np.array_equal(bits_of(view), want_dense)


@pytest.mark.parametrize("dtype", [0, 1])
@pytest.mark.parametrize("shape", [(128, 128), (130, 256), (64, 576), (512, 4608), (6, 16), (256, 2304)])
@pytest.mark.parametrize("inplace", [False, True])
def test_prune24_tile_fused_compress_both_layouts(spfy, orc, cuda, dtype, shape, inplace):
    """cols % 16 == 0 and a weight-sized matrix: TILE prune + compress is ONE kernel (prune24_tile_fused_kernel).  Its outputs must be those of
    the STRIP compressor applied to the TILE-pruned matrix, in the canonical and in the SM100 (GEMM operand) layout,
    padding tiles included, in place or not"""
    rows, cols = shape
    bits = rand_bits(orc, dtype, shape, seed=5 * rows + cols)
    bits[::5, : cols // 2] &= 0xFC00  # ties, and zeros that survive the prune
    want_dense, _ = orc.prune24_tile(dtype, bits)
    ref = orc.prune24_strip(dtype, want_dense, want_mask=False)
    ov, om = orc.pack_sm100(ref["vals"], ref["meta"], rows, cols)
    for layout in (spfy.LAYOUT_SM100, spfy.LAYOUT_CANONICAL):
        a = to_dev(bits, dtype, cuda)
        dense = a if inplace else torch.empty_like(a)
        before = spfy.launch_count()
        comp = spfy.prune24(a, out_dense=dense, mode=spfy.PRUNE_TILE_MAG, layout=layout)
        assert spfy.launch_count() - before == 1
        assert np.array_equal(bits_of(dense), want_dense)
        if layout == spfy.LAYOUT_SM100:
            assert np.array_equal(comp.vals.cpu().numpy(), ov)
            assert np.array_equal(comp.meta.cpu().numpy(), om)
        else:
            assert np.array_equal(comp.vals.cpu().numpy().view(np.uint16).reshape(rows, -1), ref["vals"])
            assert np.array_equal(comp.meta.cpu().numpy().reshape(rows, -1), ref["meta"])


def test_prune24_batched_equals_single(spfy, orc, cuda):
    import ctypes
    shapes = [(64, 147), (64, 576), (128, 1152), (256, 2304), (512, 4608), (130, 260)] * 20  # > 96 items
    ins, comps, singles = [], [], []
    for i, (r, c) in enumerate(shapes):
        a = to_dev(rand_bits(orc, 0, (r, c), seed=1000 + i), 0, cuda)
        ins.append(a)
        vb, mb = spfy.compressed_bytes(torch.float16, r, c)
        comps.append((torch.zeros(vb, dtype=torch.uint8, device=cuda), torch.zeros(mb, dtype=torch.uint8, device=cuda)))
        singles.append(spfy.prune24(a))

    class Item(ctypes.Structure):
        _fields_ = [("in_", ctypes.c_void_p), ("ld_in", ctypes.c_size_t), ("out_dense", ctypes.c_void_p),
                    ("ld_out", ctypes.c_size_t), ("comp_vals", ctypes.c_void_p), ("meta", ctypes.c_void_p),
                    ("rows", ctypes.c_size_t), ("cols", ctypes.c_size_t)]

    items = (Item * len(shapes))()
    for i, ((r, c), a, (v, m)) in enumerate(zip(shapes, ins, comps)):
        items[i] = Item(a.data_ptr(), a.stride(0), None, 0, v.data_ptr(), m.data_ptr(), r, c)
    spfy.capi.spfy_prune24_batched(spfy.F16, spfy.PRUNE_STRIP_MAG, spfy.LAYOUT_SM100, ctypes.cast(items, ctypes.c_void_p),
                                   len(shapes), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    for (v, m), s in zip(comps, singles):
        assert torch.equal(v, s.vals) and torch.equal(m, s.meta)


@pytest.mark.parametrize("dt", [0, 1])
def test_prune24_batched_tile_mode(spfy, orc, cuda, dt):
    """TILE_MAG over a whole weight set in one launch (plus the ragged first-conv matrix, which takes the general
    route on its own): pruned in place like cusparseLtSpMMAPrune(dA, dA) (spmma.hxx:86), bit-exact with the oracle's
    cusparseLt-pinned TILE selection, compressed operand identical to the single call's"""
    tdt = torch.float16 if dt == 0 else torch.bfloat16
    shapes = [(64, 147), (64, 576), (128, 1152), (256, 2304), (130, 256), (512, 512), (64, 64)] * 15  # > 96 items
    mats, comps, want_dense, singles = [], [], [], []
    for i, (r, c) in enumerate(shapes):
        bits = rand_bits(orc, dt, (r, c), seed=2000 + i)
        a = to_dev(bits, dt, cuda)
        mats.append(a)
        comps.append(spfy.alloc_compressed(tdt, r, c, cuda))
        if i < 7:
            want_dense.append(orc.prune24_tile(dt, bits)[0])
        ref = a.clone()
        singles.append(spfy.prune24(ref, inplace=True, mode=spfy.PRUNE_TILE_MAG))
    before = spfy.launch_count()
    spfy.prune24_batched(mats, comps, mode=spfy.PRUNE_TILE_MAG)
    torch.cuda.synchronize()
    # 90 eligible matrices -> one batched launch; 15 ragged ones -> tile + compress kernels each
    assert spfy.launch_count() - before == 1 + 15 * 2
    for i, want in enumerate(want_dense):
        assert np.array_equal(bits_of(mats[i]), want), shapes[i]
    for c, s_ in zip(comps, singles):
        assert torch.equal(c.vals, s_.vals) and torch.equal(c.meta, s_.meta)


# ------------------------------------------------------------------ A4 spmma
SPMMA_SHAPES = [  # (M, K, N)
    (128, 128, 128), (128, 256, 256), (64, 64, 512), (64, 147, 1024), (64, 148, 1000), (256, 512, 384),
    (512, 1152, 520), (130, 260, 264), (2048, 512, 392), (16, 32, 8), (128, 4608, 128),
]


@pytest.mark.parametrize("M,K,N", SPMMA_SHAPES)
@pytest.mark.parametrize("dt", [0, 1])
def test_spmma_matches_fp64_oracle(spfy, orc, cuda, M, K, N, dt):
    a_bits = rand_bits(orc, dt, (M, K), seed=M + 3 * K + 7 * N)
    b_bits = rand_bits(orc, dt, (K, N), seed=M + 3 * K + 7 * N + 1)
    pr = orc.prune24_strip(dt, a_bits, want_mask=False)
    want = orc.spmma_f64(dt, pr["dense"], b_bits)
    comp = spfy.prune24(to_dev(a_bits, dt, cuda))
    d = spfy.spmma_compressed(comp, to_dev(b_bits, dt, cuda))
    got = d.float().cpu().numpy().astype(np.float64)
    assert rel_err(got, want) <= REL_TOL


# one shape per launch class of the v2 kernel (single-problem entry point):
#   resident A (whole compressed A in shared memory, N >= 148 tiles; three ring geometries), streaming G=2 (two m-tiles share a
#   B slice), streaming G=1, m-group slices of A resident (A too large as a whole, >= 2 waves of units); ragged M / K / N
#   edges in each
CLASS_SHAPES = [(64, 147, 19008), (200, 72, 19080), (512, 128, 18944), (100, 576, 19000),  # resident
                (64, 64, 19008), (256, 64, 19200), (200, 40, 19080), (130, 24, 18960),        # resident, k <= 64 (half stages)
                (128, 512, 19008),                                                             # resident, large operand
                (512, 200, 9600), (300, 264, 9480), (1024, 256, 4800),                       # stream, G=2
                (1024, 256, 9600), (400, 256, 19000), (384, 250, 19000),                     # resident m-group slices (G=2; short / odd last group)
                (128, 1152, 2048), (96, 2304, 1000),                                         # stream, G=1
                (256, 384, 25088), (384, 640, 12752), (512, 600, 11480)]                      # stream, G=2 with a split tail wave


@pytest.mark.parametrize("M,K,N", CLASS_SHAPES)
def test_spmma_every_launch_class(spfy, orc, cuda, M, K, N):
    a_bits = rand_bits(orc, 0, (M, K), seed=M * 3 + K)
    b_bits = rand_bits(orc, 0, (K, N), seed=N)
    pr = orc.prune24_strip(0, a_bits, want_mask=False)
    want = orc.spmma_f64(0, pr["dense"], b_bits)
    comp = spfy.prune24(to_dev(a_bits, 0, cuda))
    d = spfy.spmma_compressed(comp, to_dev(b_bits, 0, cuda))
    assert rel_err(d.float().cpu().numpy().astype(np.float64), want) <= REL_TOL


def test_spmma_plan_matches_single_calls_and_oracle(spfy, orc, cuda):
    """grouped persistent launch over a mixed list (all three classes, both opB, alpha/beta, bf16 is a
    separate plan): every output must equal the single-call result bit for bit and the oracle within tol."""
    shapes = [(64, 147, 19008), (512, 128, 18944), (256, 64, 20000), (512, 200, 1600), (1024, 256, 1200),
              (128, 1152, 520), (256, 2304, 392), (2048, 512, 264), (64, 576, 19000), (130, 260, 264),
              (1024, 256, 9600), (384, 256, 19000), (384, 640, 12752)]  # the last one: single call splits its tail wave
    problems, singles, wants = [], [], []
    for i, (M, K, N) in enumerate(shapes):
        a_bits = rand_bits(orc, 0, (M, K), seed=500 + i)
        op_t = i % 4 == 3
        b_bits = rand_bits(orc, 0, (N, K) if op_t else (K, N), seed=600 + i)
        alpha, beta = (0.5, 0.25) if i % 3 == 1 else (1.0, 0.0)
        c_bits = rand_bits(orc, 0, (M, N), seed=700 + i) if beta else None
        pr = orc.prune24_strip(0, a_bits, want_mask=False)
        wants.append(orc.spmma_f64(0, pr["dense"], b_bits, c_bits=c_bits, alpha=alpha, beta=beta, op_b=int(op_t)))
        comp = spfy.prune24(to_dev(a_bits, 0, cuda))
        b = to_dev(b_bits, 0, cuda)
        c = to_dev(c_bits, 0, cuda) if beta else None
        op_b = spfy.OP_T if op_t else spfy.OP_N
        singles.append(spfy.spmma_compressed(comp, b, c=c, alpha=alpha, beta=beta, op_b=op_b))
        problems.append(dict(comp=comp, b=b, c=c, out=torch.zeros(M, N, dtype=torch.float16, device=cuda), alpha=alpha,
                             beta=beta, op_b=op_b))
    plan = spfy.SpmmaPlan(problems)
    assert 1 <= plan.launches <= 12  # one per (launch class, opB) present
    before = spfy.launch_count()
    plan.run()
    plan.run()  # a plan is reusable
    torch.cuda.synchronize()
    assert spfy.launch_count() - before == 2 * plan.launches
    for q, single, want in zip(problems, singles, wants):
        assert torch.equal(q["out"], single)
        assert rel_err(q["out"].float().cpu().numpy().astype(np.float64), want) <= REL_TOL
    plan.close()


@pytest.mark.parametrize("M,K,N", [(64, 147, 19008), (256, 64, 19200), (512, 200, 9600), (1024, 256, 9600), (128, 1152, 2048),
                                   (104, 576, 19000), (136, 260, 264), (2048, 512, 392)])
@pytest.mark.parametrize("dt", [0, 1])
def test_spmma_transposed_output(spfy, orc, cuda, M, K, N, dt):
    """SPFY_OUT_T: D written as [n][m] (NHWC for a convolution layer) -- bit for bit the transpose of the row-major
    result in every launch class, with alpha, a pitched destination and rows that do not fill a 32-row store box"""
    comp = spfy.prune24(to_dev(rand_bits(orc, dt, (M, K), seed=M + K), dt, cuda))
    b = to_dev(rand_bits(orc, dt, (K, N), seed=N + 1), dt, cuda)
    want = spfy.spmma_compressed(comp, b, alpha=0.5)
    ld = (M + 15) // 8 * 8
    buf = torch.full((N, ld), 7.0, dtype=want.dtype, device=cuda)
    spfy.spmma_compressed(comp, b, alpha=0.5, out=buf[:, :M], out_t=True)
    torch.cuda.synchronize()
    assert torch.equal(buf[:, :M], want.t())
    assert bool((buf[:, M:] == 7.0).all())  # nothing written beyond column m
    if K % 8 == 0:  # opB = T needs k % 8 == 0
        buf2 = torch.empty((N, ld), dtype=want.dtype, device=cuda)
        spfy.spmma_compressed(comp, b.t().contiguous(), op_b=spfy.OP_T, out=buf2[:, :M], out_t=True)
        assert torch.equal(buf2[:, :M], spfy.spmma_compressed(comp, b).t())
    with pytest.raises(spfy.SpfyError):
        spfy.spmma_compressed(comp, b, c=want, beta=1.0, out_t=True)
    if M == 104:  # the contiguous dimension of an output is a multiple of 8 elements
        odd = spfy.prune24(to_dev(rand_bits(orc, dt, (100, K), seed=1), dt, cuda))
        with pytest.raises(spfy.SpfyError) as e:
            spfy.spmma_compressed(odd, b, out=buf[:, :100], out_t=True)
        assert e.value.code == spfy.capi.E_UNSUPPORTED


def test_spmma_plan_mixes_output_layouts(spfy, orc, cuda):
    shapes = [(64, 147, 19008), (256, 64, 19200), (1024, 256, 9600), (256, 2304, 392), (136, 260, 264), (512, 128, 18944)]
    problems, wants = [], []
    for i, (M, K, N) in enumerate(shapes):
        comp = spfy.prune24(to_dev(rand_bits(orc, 0, (M, K), seed=40 + i), 0, cuda))
        b = to_dev(rand_bits(orc, 0, (K, N), seed=60 + i), 0, cuda)
        t = i % 2 == 0
        wants.append(spfy.spmma_compressed(comp, b))
        ld = (M + 7) // 8 * 8  # the pitch of a transposed output is a multiple of 8 elements like every other
        out = torch.zeros(N, ld, dtype=torch.float16, device=cuda)[:, :M] if t else torch.zeros(M, N, dtype=torch.float16, device=cuda)
        problems.append(dict(comp=comp, b=b, out=out, out_t=t))
    plan = spfy.SpmmaPlan(problems)
    plan.run()
    torch.cuda.synchronize()
    for q, want in zip(problems, wants):
        assert torch.equal(q["out"], want.t() if q.get("out_t") else want)
    plan.close()


def test_spmma_plan_replicated_outputs(spfy, orc, cuda):
    """spfy_spmma_plan_create_replicated on one GPU: every replica (here: other buffers of the same device; across GPUs:
    peer mappings, tests/mg_worker.py) ends up bit-identical to the primary output, for every launch class, both
    epilogue paths and a pitched destination."""
    shapes = [(64, 147, 19008), (256, 64, 19200), (512, 200, 1600), (1024, 256, 9600), (128, 1152, 520), (130, 260, 264)]
    problems, copies = [], []
    for i, (M, K, N) in enumerate(shapes):
        comp = spfy.prune24(to_dev(rand_bits(orc, 0, (M, K), seed=900 + i), 0, cuda))
        b = to_dev(rand_bits(orc, 0, (K, N), seed=950 + i), 0, cuda)
        beta = 0.5 if i % 2 else 0.0
        c = to_dev(rand_bits(orc, 0, (M, N), seed=970 + i), 0, cuda) if beta else None
        ld = N + 8 * (i % 3)  # all copies of a problem share the row pitch
        bufs = [torch.zeros(M, ld, dtype=torch.float16, device=cuda) for _ in range(3)]
        copies.append((N, bufs))
        problems.append(dict(comp=comp, b=b, c=c, beta=beta, out=bufs[0][:, :N], replicas=[t.data_ptr() for t in bufs[1:]]))
    plan = spfy.SpmmaPlan(problems)
    assert plan.replicas == 2
    plan.run()
    torch.cuda.synchronize()
    for q, (N, bufs) in zip(problems, copies):
        single = spfy.spmma_compressed(q["comp"], q["b"], c=q["c"], beta=q["beta"])
        assert torch.equal(bufs[0][:, :N], single)
        for t in bufs[1:]:
            assert torch.equal(t, bufs[0])  # the padding columns stay zero in every copy
    plan.close()
    with pytest.raises(ValueError):
        spfy.SpmmaPlan([dict(problems[0], replicas=[copies[0][1][1].data_ptr()]), problems[1]])
    with pytest.raises(spfy.SpfyError):
        spfy.SpmmaPlan([dict(problems[0], replicas=[copies[0][1][1].data_ptr() + 2])])


def test_spmma_plan_resnet18_table(spfy, cuda):
    """BASELINE config 1 end to end at reduced batch: every layer of datasets/resnet18.csv through the
    batched prune + plan path equals a dense matmul of the pruned weights."""
    torch.manual_seed(1)
    gemms = [spfy.shapes.to_gemm(s, "weights", 2) for s in spfy.shapes.read_shapes("resnet18.csv")]
    ws = [(torch.rand(g.M, g.K, device=cuda) * 2 - 1).half() for g in gemms]
    comps = [spfy.alloc_compressed(torch.float16, g.M, g.K, cuda) for g in gemms]
    spfy.prune24_batched(ws, comps)
    problems = []
    for g, comp in zip(gemms, comps):
        b = torch.randint(-2, 3, (g.K, g.N), device=cuda).half()
        problems.append(dict(comp=comp, b=b, out=torch.empty(g.M, g.N, dtype=torch.float16, device=cuda)))
    plan = spfy.SpmmaPlan(problems)
    plan.run()
    torch.cuda.synchronize()
    for w, q in zip(ws, problems):
        wd = torch.empty_like(w)
        spfy.prune24(w, out_dense=wd, compress=False)
        ref = wd.float() @ q["b"].float()
        assert float((q["out"].float() - ref).abs().max() / ref.abs().max()) < 2e-3


@pytest.mark.parametrize("M,K,N", [(128, 256, 256), (64, 147 + 5, 520), (256, 1024, 136), (64, 64, 19008),
                                   (128, 48, 19200), (64, 256, 19008)])
def test_spmma_transposed_b(spfy, orc, cuda, M, K, N):
    a_bits = rand_bits(orc, 0, (M, K), seed=91)
    bt_bits = rand_bits(orc, 0, (N, K), seed=92)
    pr = orc.prune24_strip(0, a_bits, want_mask=False)
    want = orc.spmma_f64(0, pr["dense"], bt_bits, op_b=1)
    comp = spfy.prune24(to_dev(a_bits, 0, cuda))
    d = spfy.spmma_compressed(comp, to_dev(bt_bits, 0, cuda), op_b=spfy.OP_T)
    assert rel_err(d.float().cpu().numpy().astype(np.float64), want) <= REL_TOL


def test_spmma_alpha_beta(spfy, orc, cuda):
    M, K, N = 192, 320, 264
    a_bits = rand_bits(orc, 0, (M, K), seed=1)
    b_bits = rand_bits(orc, 0, (K, N), seed=2)
    c_bits = rand_bits(orc, 0, (M, N), seed=3)
    pr = orc.prune24_strip(0, a_bits, want_mask=False)
    want = orc.spmma_f64(0, pr["dense"], b_bits, c_bits=c_bits, alpha=0.5, beta=-1.5)
    comp = spfy.prune24(to_dev(a_bits, 0, cuda))
    c = to_dev(c_bits, 0, cuda)
    d = spfy.spmma_compressed(comp, to_dev(b_bits, 0, cuda), c=c, out=c, alpha=0.5, beta=-1.5)  # D aliases C
    assert rel_err(d.float().cpu().numpy().astype(np.float64), want) <= REL_TOL


def test_spmma_exact_on_small_integers(spfy, orc, cuda):
    """integer-valued inputs make the fp32 accumulation exact: the output must then EQUAL the oracle,
    which pins the metadata semantics (a wrong index would pick a different B row)."""
    rng = np.random.default_rng(5)
    M, K, N = 256, 384, 256
    a = rng.integers(-4, 5, (M, K)).astype(np.float32)
    b = rng.integers(-4, 5, (K, N)).astype(np.float32)
    a_bits, b_bits = orc.from_f32(0, a), orc.from_f32(0, b)
    pr = orc.prune24_strip(0, a_bits, want_mask=False)
    want = orc.spmma_f64(0, pr["dense"], b_bits)
    comp = spfy.prune24(to_dev(a_bits, 0, cuda))
    d = spfy.spmma_compressed(comp, to_dev(b_bits, 0, cuda))
    assert np.array_equal(d.float().cpu().numpy().astype(np.float64), want)


def test_spmma_linearity_full_size(spfy, cuda):
    """BASELINE-size property check (ResNet-50 layer 196,512,4608,32 in the weights orientation):
    the kernel is linear in B, and equals a dense matmul of the pruned weights."""
    torch.manual_seed(0)
    M, K, N = 512, 4608, 6272
    w = (torch.rand(M, K, device=cuda) * 2 - 1).half()
    wd = torch.empty_like(w)
    comp = spfy.prune24(w, out_dense=wd)
    b1 = torch.randint(-2, 3, (K, N), device=cuda).half()
    b2 = torch.randint(-2, 3, (K, N), device=cuda).half()
    d1 = spfy.spmma_compressed(comp, b1).float()
    d2 = spfy.spmma_compressed(comp, b2).float()
    d12 = spfy.spmma_compressed(comp, b1 + b2).float()
    ref = wd.float() @ (b1 + b2).float()
    scale = ref.abs().max()
    assert float((d12 - ref).abs().max() / scale) < 2e-3
    assert float((d1 + d2 - d12).abs().max() / scale) < 4e-3


def test_spmma_reference_style_call(spfy, orc, cuda):
    """the reference-shaped entry point: prune in place, compress, multiply into C; three timings"""
    m, n, k = 128, 256, 512
    a_bits = rand_bits(orc, 0, (m, k), seed=8)
    b_bits = rand_bits(orc, 0, (k, n), seed=9)
    b = to_dev(b_bits, 0, cuda).reshape(-1)
    tile_dense, _ = orc.prune24_tile(0, a_bits)
    strip_dense = orc.prune24_strip(0, a_bits, want_mask=False)["dense"]
    # default: TILE, what the reference asks cusparseLt for (spmma.hxx:86); then the per-row STRIP variant
    for kwargs, pruned in (({}, tile_dense), ({"prune_mode": spfy.PRUNE_STRIP_MAG}, strip_dense)):
        a = to_dev(a_bits, 0, cuda).reshape(-1)
        c = torch.zeros(m * n, dtype=torch.float16, device=cuda)
        times = spfy.spmma(a, b, c, m, n, k, 1, **kwargs)
        assert len(times) == 3 and all(t >= 0 for t in times)
        assert np.array_equal(bits_of(a).reshape(m, k), pruned)  # A pruned in place
        want = orc.spmma_f64(0, pruned, b_bits)
        assert rel_err(c.view(m, n).float().cpu().numpy().astype(np.float64), want) <= REL_TOL


def test_packed_container_feeds_spmma(spfy, orc, cuda):
    """prune + compress -> container bytes -> back onto the device: the reloaded operand is the same image and gives
    the same product"""
    m, n, k = 256, 320, 1152
    a = to_dev(rand_bits(orc, 0, (m, k), seed=41), 0, cuda)
    b = to_dev(rand_bits(orc, 0, (k, n), seed=42), 0, cuda)
    comp = spfy.prune24(a, mode=spfy.PRUNE_TILE_MAG, out_dense=torch.empty_like(a))
    blob = spfy.pack_compressed(comp)
    back = spfy.unpack_compressed(blob, cuda)
    assert (back.rows, back.cols, back.dtype, back.layout) == (m, k, torch.float16, spfy.LAYOUT_SM100)
    assert torch.equal(back.vals, comp.vals) and torch.equal(back.meta, comp.meta)
    assert torch.equal(spfy.spmma_compressed(back, b), spfy.spmma_compressed(comp, b))
    with pytest.raises(spfy.SpfyError):
        spfy.unpack_compressed(blob[:100] + bytes([blob[100] ^ 1]) + blob[101:], cuda)


def test_spmma_rejects_bad_arguments(spfy, cuda):
    a = torch.zeros(128, 128, dtype=torch.float16, device=cuda)
    comp = spfy.prune24(a)
    b = torch.zeros(128, 12, dtype=torch.float16, device=cuda)  # n = 12 is not a multiple of 8
    with pytest.raises(spfy.SpfyError) as e:
        spfy.spmma_compressed(comp, b)
    assert e.value.code == spfy.capi.E_UNSUPPORTED


# ------------------------------------------------------------------ unstructured path
@pytest.mark.parametrize("rows,cols", [(64, 147), (128, 1152), (1, 1), (300, 37), (256, 2304)])
@pytest.mark.parametrize("sparsity", [0.5, 0.9, 0.95])
@pytest.mark.parametrize("tdt,code", [(torch.float32, 2), (torch.float16, 0)])
def test_threshold_to_coo_bit_exact(spfy, orc, cuda, rows, cols, sparsity, tdt, code):
    rng = np.random.default_rng(rows * 13 + cols)
    w = rng.uniform(-1, 1, (rows, cols)).astype(np.float32)
    if code == 0:
        wb = orc.from_f32(0, w)
        w = orc.to_f32(0, wb)
        a = to_dev(wb, 0, cuda)
        host = wb
    else:
        a = torch.from_numpy(w).to(cuda)
        host = w
    thr = float(np.quantile(np.abs(w), sparsity))
    ri, ci, va, rp = orc.threshold_to_coo(code, host, thr)
    gri, gci, gva, nnz, grp = spfy.threshold_to_coo(a, thr, want_csr=True)
    assert nnz == ri.size
    assert np.array_equal(gri.cpu().numpy(), ri) and np.array_equal(gci.cpu().numpy(), ci)
    assert np.array_equal(gva.cpu().numpy(), va)
    assert np.array_equal(grp.cpu().numpy(), rp)
    assert np.array_equal(spfy.coo_to_csr(gri, rows).cpu().numpy(), orc.coo_to_csr(ri, rows))


@pytest.mark.parametrize("rows,cols,tdt", [(4096, 4608, torch.float32), (12544, 147, torch.float32),
                                            (8192, 2304, torch.float16), (1, 70000, torch.float32)])
@pytest.mark.parametrize("keep", [0.5, 0.02])
def test_threshold_full_size_properties(spfy, cuda, rows, cols, tdt, keep):
    """sizes the CPU oracle would take too long on: the COO must be the masked input itself (scatter it back),
    sorted by (row, column), with nnz and row_ptr that agree with a count done by torch -- many chunks of the
    chained scan, rows shorter and longer than a chunk, and the 16-bit / unaligned load paths"""
    g = torch.Generator(device="cuda").manual_seed(rows + cols)
    a = (torch.rand(rows, cols, device=cuda, generator=g) * 2 - 1).to(tdt)
    thr = 1.0 - keep
    ri, ci, va, nnz, rp = spfy.threshold_to_coo(a, thr, want_csr=True)
    mask = a.float().abs() > thr
    assert nnz == int(mask.sum())
    key = ri.long() * cols + ci.long()
    assert bool((key[1:] > key[:-1]).all())
    back = torch.zeros(rows, cols, dtype=torch.float32, device=cuda)
    back[ri.long(), ci.long()] = va
    assert torch.equal(back, torch.where(mask, a.float(), torch.zeros((), device=cuda)))
    want_rp = torch.zeros(rows + 1, dtype=torch.int64, device=cuda)
    want_rp[1:] = mask.sum(1).cumsum(0)
    assert torch.equal(rp.long(), want_rp)


def test_prune24_tile_full_size_properties(spfy, cuda):
    """TILE on the largest ResNet weight matrix (512 x 4608) and a 4096 x 4608 one: two kept per row AND per
    column of every 4x4 tile, kept entries untouched, idempotent, and never lighter than any fixed pattern"""
    for rows, cols in ((512, 4608), (4096, 4608)):
        g = torch.Generator(device="cuda").manual_seed(rows)
        a = (torch.rand(rows, cols, device=cuda, generator=g) * 2 - 1).half()
        d1 = torch.empty_like(a)
        spfy.prune24(a, out_dense=d1, mode=spfy.PRUNE_TILE_MAG, compress=False)
        t = (d1 != 0).view(rows // 4, 4, cols // 4, 4)
        assert int(t.sum(3).max()) <= 2 and int(t.sum(1).max()) <= 2
        kept = d1 != 0
        assert torch.equal(d1[kept], a[kept])
        d2 = torch.empty_like(a)
        spfy.prune24(d1, out_dense=d2, mode=spfy.PRUNE_TILE_MAG, compress=False)
        assert torch.equal(d1, d2)
        mass = d1.float().abs().view(rows // 4, 4, cols // 4, 4).sum((1, 3))
        checker = torch.tensor([[1, 1, 0, 0], [0, 0, 1, 1], [1, 1, 0, 0], [0, 0, 1, 1]], device=cuda, dtype=torch.float32)
        fixed = (a.float().abs().view(rows // 4, 4, cols // 4, 4) * checker.view(1, 4, 1, 4)).sum((1, 3))
        assert bool((mass >= fixed - 1e-3).all())
        assert spfy.prune24_check(d1) == 0


@pytest.mark.parametrize("m,k,n,nb,sparsity", [(64, 147, 96, 2, 0.5), (128, 576, 200, 3, 0.9), (256, 1152, 49, 4, 0.95),
                                              (33, 70, 17, 1, 0.5), (512, 512, 64, 2, 0.9)])
def test_batched_coo_spmm_matches_oracle(spfy, orc, cuda, m, k, n, nb, sparsity):
    rng = np.random.default_rng(m + k + n)
    w = rng.uniform(-1, 1, (m, k)).astype(np.float32)
    thr = float(np.quantile(np.abs(w), sparsity))
    ri, ci, va, _ = orc.threshold_to_coo(2, w, thr)
    B = rng.uniform(-1, 1, (nb, n, k)).astype(np.float32)
    C0 = rng.uniform(-1, 1, (nb, n, m)).astype(np.float32)
    for alpha, beta in [(1.0, 0.0), (0.75, 0.5)]:
        want = orc.spmm_coo_batched_f64(m, k, n, nb, ri, ci, va, B, C0, alpha, beta)
        dC = torch.from_numpy(C0.copy()).to(cuda)
        ms = spfy.batched.strided_coo(m, k, ri.size, k, n, nb, torch.from_numpy(ri).to(cuda),
                                      torch.from_numpy(ci).to(cuda), torch.from_numpy(va).to(cuda),
                                      torch.from_numpy(B).to(cuda), dC, alpha=alpha, beta=beta)
        assert ms >= 0
        assert np.allclose(dC.cpu().numpy(), want, rtol=2e-4, atol=2e-4)


def test_batched_coo_spmm_empty_rows_and_unsorted_columns(spfy, orc, cuda):
    m, k, n, nb = 40, 64, 24, 2
    ri = np.array([0, 0, 0, 5, 5, 39], dtype=np.int32)
    ci = np.array([63, 1, 30, 7, 7, 0], dtype=np.int32)  # unsorted within a row + a duplicate entry
    va = np.array([1.0, -2.0, 0.5, 3.0, 1.0, -1.0], dtype=np.float32)
    rng = np.random.default_rng(0)
    B = rng.uniform(-1, 1, (nb, n, k)).astype(np.float32)
    want = orc.spmm_coo_batched_f64(m, k, n, nb, ri, ci, va, B)
    dC = torch.full((nb, n, m), 9.0, dtype=torch.float32, device=cuda)
    spfy.batched.strided_coo(m, k, ri.size, k, n, nb, torch.from_numpy(ri).to(cuda), torch.from_numpy(ci).to(cuda),
                             torch.from_numpy(va).to(cuda), torch.from_numpy(B).to(cuda), dC)
    assert np.allclose(dC.cpu().numpy(), want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("shuffle", [False, True])
def test_batched_coo_spmm_many_chunks(spfy, orc, cuda, shuffle):
    """K spans several staged chunks and rows hold more than one 32-entry request; with shuffled
    columns inside every row the kernel must fall back from its per-row cursor to rescanning."""
    m, k, n, nb = 96, 960, 150, 2
    rng = np.random.default_rng(11)
    w = rng.uniform(-1, 1, (m, k)).astype(np.float32)
    w[5] = 0.0   # an empty row
    w[7, :] = rng.uniform(1, 2, k)  # a dense row
    thr = float(np.quantile(np.abs(w), 0.8))
    ri, ci, va, _ = orc.threshold_to_coo(2, w, thr)
    if shuffle:
        for r in range(m):
            sel = np.nonzero(ri == r)[0]
            p = rng.permutation(sel.size)
            ci[sel], va[sel] = ci[sel][p], va[sel][p]
    B = rng.uniform(-1, 1, (nb, n, k)).astype(np.float32)
    want = orc.spmm_coo_batched_f64(m, k, n, nb, ri, ci, va, B)
    dC = torch.full((nb, n, m), 7.0, dtype=torch.float32, device=cuda)
    spfy.batched.strided_coo(m, k, ri.size, k, n, nb, torch.from_numpy(ri).to(cuda), torch.from_numpy(ci).to(cuda),
                             torch.from_numpy(va).to(cuda), torch.from_numpy(B).to(cuda), dC)
    assert np.allclose(dC.cpu().numpy(), want, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("tdt", [torch.float32, torch.float16])
def test_blocked_ell_spmm_matches_oracle(spfy, orc, cuda, tdt):
    """the reference driver's construction (examples/spmm.cu:45-84): block 2, ell_cols = k/2,
    values 1.., sorted unique block-column ids per block row."""
    m, n, k, nb, block = 64, 48, 128, 3, 2
    ell_cols = k // 2
    bcols = ell_cols // block
    rng = np.random.default_rng(4)
    B = (rng.integers(-3, 4, (n, k))).astype(np.float32)
    cis, vas, cs, wants = [], [], [], []
    for b in range(nb):
        ci = np.stack([np.sort(rng.choice(k // block, bcols, replace=False)) for _ in range(m // block)]).astype(np.int64)
        va = ((np.arange(m * ell_cols) % 7) - 3).astype(np.float32).reshape(m, ell_cols)
        wants.append(orc.spmm_bell_f64(m, k, n, block, ell_cols, ci, va, B))
        cis.append(torch.from_numpy(ci).to(cuda))
        vas.append(torch.from_numpy(va).to(cuda).to(tdt))
        cs.append(torch.zeros(n, m, dtype=tdt, device=cuda))
    ms = spfy.batched.spmm(cis, vas, torch.from_numpy(B).to(cuda).to(tdt), cs, m, n, k, block, ell_cols)
    assert ms >= 0
    for c, want in zip(cs, wants):
        assert np.array_equal(c.float().cpu().numpy().astype(np.float64), want)


@pytest.mark.parametrize("order", ["sorted", "shuffled", "padded"])
def test_blocked_ell_spmm_multi_chunk(spfy, orc, cuda, order):
    """fp32 blocked-ELL through the shared-memory kernel: several staged chunks of k, 128-row tiles, ragged
    n; block-column ids ascending (cursor mode), shuffled, or with padding slots (< 0) in the middle
    (both fall back to rescanning every row per chunk)."""
    m, n, k, nb, block = 200, 150, 448, 2, 4
    ell_cols = k // 2
    bcols = ell_cols // block
    rng = np.random.default_rng(21)
    B = rng.uniform(-1, 1, (n, k)).astype(np.float32)
    cis, vas, cs, wants = [], [], [], []
    for b in range(nb):
        ci = np.stack([np.sort(rng.choice(k // block, bcols, replace=False)) for _ in range(m // block)]).astype(np.int64)
        if order == "shuffled":
            ci = np.stack([rng.permutation(r) for r in ci])
        elif order == "padded":
            ci[:, ::5] = -1
        va = rng.uniform(-1, 1, (m, ell_cols)).astype(np.float32)
        wants.append(orc.spmm_bell_f64(m, k, n, block, ell_cols, ci, va, B))
        cis.append(torch.from_numpy(ci).to(cuda))
        vas.append(torch.from_numpy(va).to(cuda))
        cs.append(torch.full((n, m), 3.0, dtype=torch.float32, device=cuda))
    spfy.batched.spmm(cis, vas, torch.from_numpy(B).to(cuda), cs, m, n, k, block, ell_cols)
    for c, want in zip(cs, wants):
        assert np.allclose(c.cpu().numpy().astype(np.float64), want, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("density", [0.6, 0.05])
@pytest.mark.parametrize("m,k,n,nb", [(96, 250, 131, 2), (200, 97, 64, 3), (64, 1000, 40, 1)])
def test_csr_entry_picks_its_kernel_on_the_device(spfy, orc, cuda, m, k, n, nb, density):
    """the CSR entry has no host-side nnz: both SpMM kernels are launched and a device flag lets exactly one
    run (dense walk from 35 % non-zeros).  k is not a multiple of the 96-wide chunk (nor of 4), rows are
    shuffled (rescan mode) and one entry is duplicated (entries of one cell add up, like cuSPARSE COO)."""
    rng = np.random.default_rng(m * k + n)
    w = rng.uniform(-1, 1, (m, k)).astype(np.float32)
    ri, ci, va, _ = orc.threshold_to_coo(2, w, float(np.quantile(np.abs(w), 1.0 - density)))
    ri, ci, va = (np.concatenate([x, x[-1:]]) for x in (ri, ci, va))  # duplicate of the last entry
    for r in range(0, m, 3):
        sel = np.nonzero(ri == r)[0]
        pm = rng.permutation(sel.size)
        ci[sel], va[sel] = ci[sel][pm], va[sel][pm]
    B = rng.uniform(-1, 1, (nb, n, k)).astype(np.float32)
    C0 = rng.uniform(-1, 1, (nb, n, m)).astype(np.float32)
    want = orc.spmm_coo_batched_f64(m, k, n, nb, ri, ci, va, B, C0, 0.5, 2.0)
    dri = torch.from_numpy(ri).to(cuda)
    rp = spfy.coo_to_csr(dri, m)
    dC = torch.from_numpy(C0.copy()).to(cuda)
    spfy.batched.csr(m, k, n, nb, rp, torch.from_numpy(ci).to(cuda), torch.from_numpy(va).to(cuda),
                     torch.from_numpy(B).to(cuda), dC, alpha=0.5, beta=2.0)
    assert np.allclose(dC.cpu().numpy(), want, rtol=2e-4, atol=2e-4)


def test_dense_walk_unaligned_b_and_exact_integers(spfy, orc, cuda):
    """B columns that are not 16-byte aligned take the scalar staging path; small integers make every sum
    exact, so the result must equal the fp64 oracle bit for bit"""
    m, k, n, nb = 130, 333, 70, 2
    rng = np.random.default_rng(5)
    w = rng.integers(-4, 5, (m, k)).astype(np.float32)
    ri, ci, va, _ = orc.threshold_to_coo(2, w, 1.5)  # ~ 55 % non-zeros
    B = rng.integers(-3, 4, (nb, n, k)).astype(np.float32)
    want = orc.spmm_coo_batched_f64(m, k, n, nb, ri, ci, va, B)
    slab = torch.zeros(nb * n * k + 1, dtype=torch.float32, device=cuda)
    db = slab[1:]  # 4 bytes off a 16-byte boundary
    db.copy_(torch.from_numpy(B).to(cuda).view(-1))
    dC = torch.full((nb, n, m), 7.0, dtype=torch.float32, device=cuda)
    spfy.batched.strided_coo(m, k, ri.size, k, n, nb, torch.from_numpy(ri).to(cuda), torch.from_numpy(ci).to(cuda),
                             torch.from_numpy(va).to(cuda), db, dC)
    assert np.array_equal(dC.cpu().numpy().astype(np.float64), want)


@pytest.mark.parametrize("block", [2, 3])
def test_blocked_ell_sparse_rows(spfy, orc, cuda, block):
    """ell_cols = k / 8: the (row pair) x (column pair) walk for even blocks, the per-slot walk for odd ones"""
    m, n, k, nb = 132, 100, 24 * 8 * block, 2
    ell_cols = k // 8
    bcols = ell_cols // block
    rng = np.random.default_rng(31 + block)
    B = rng.uniform(-1, 1, (n, k)).astype(np.float32)
    cis, vas, cs, wants = [], [], [], []
    for b in range(nb):
        ci = np.stack([np.sort(rng.choice(k // block, bcols, replace=False)) for _ in range(m // block)]).astype(np.int64)
        va = rng.uniform(-1, 1, (m, ell_cols)).astype(np.float32)
        wants.append(orc.spmm_bell_f64(m, k, n, block, ell_cols, ci, va, B))
        cis.append(torch.from_numpy(ci).to(cuda))
        vas.append(torch.from_numpy(va).to(cuda))
        cs.append(torch.zeros(n, m, dtype=torch.float32, device=cuda))
    spfy.batched.spmm(cis, vas, torch.from_numpy(B).to(cuda), cs, m, n, k, block, ell_cols)
    for c, want in zip(cs, wants):
        assert np.allclose(c.cpu().numpy().astype(np.float64), want, rtol=2e-4, atol=2e-4)


# (the cuSPARSE golden fixtures are checked for every SpMM route in tests/test_gpu_tensor.py)


def test_launch_counter_moves(spfy, cuda):
    before = spfy.launch_count()
    a = torch.zeros(128, 128, dtype=torch.float16, device=cuda)
    spfy.prune24(a)
    assert spfy.launch_count() == before + 1
