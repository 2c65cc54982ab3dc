"""Oracle pinned against the REFERENCE's own outputs (fixtures generated on a B200 by
tests/golden/make_golden.py; nothing here needs a GPU or /root/reference).

  ref_sparsify_*.npz : the reference's sparsifyme::sparsify<BLK_M,BLK_N> itself
                       (include/sparsify.me/sparsify.hxx:24-82)  -> orc_prune_blocks_ref must be bit-exact.
  cusparselt_*.npz   : the closed library behind the reference's spmma (spmma.hxx:86-113), v0.7.1
                       -> orc_prune24_strip must be bit-exact with PRUNE_SPMMA_STRIP (incl. the tie-break);
                          orc_prune24_tile must be bit-exact with PRUNE_SPMMA_TILE (what spmma.hxx:86 requests);
                          the fp64 GEMM oracle must agree with cusparseLtMatmul within fp16 rounding.
  tile_*.npz         : cusparseLt 0.7.1 PRUNE_SPMMA_TILE on the probes of tests/golden/make_tile_probe.py: every
                       face of the 4x4 pattern polytope (= every possible exact tie), every 2-level subset
                       weighting, tie-rich small integers, and wide-range inputs whose fp32 sums round
                       -> orc_prune24_tile must pick the library's pattern on every tile.
  cusparse_coo_*.npz, cusparse_bell_*.npz : cuSPARSE 12.x driven with the reference's call sequences for
                       batched::strided_coo (spmm.hxx:164-187, COO_ALG4) and batched::spmm (blocked-ELL,
                       :57-67,107-110).  Inputs are multiples of 1/64, so every sum is exact in fp32 and the
                       fp64 oracles must reproduce the library's output BIT FOR BIT.
"""
import glob
import hashlib
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, GOLD)
from make_golden import gen, gen_f32  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


SPARSIFY = sorted(glob.glob(os.path.join(GOLD, "ref_sparsify_*.npz")))
CUSPLT = sorted(glob.glob(os.path.join(GOLD, "cusparselt_*.npz")))


CUSP_COO = sorted(glob.glob(os.path.join(GOLD, "cusparse_coo_*.npz")))
CUSP_BELL = sorted(glob.glob(os.path.join(GOLD, "cusparse_bell_*.npz")))


def test_fixtures_present():
    assert len(SPARSIFY) >= 12 and len(CUSPLT) >= 5 and len(CUSP_COO) >= 4 and len(CUSP_BELL) >= 3


def coo_case(z):
    """inputs of a cusparse_coo fixture, regenerated (same generator as oracle/cusparse_ref.cu)"""
    m, k, n, nb = int(z["m"]), int(z["k"]), int(z["n"]), int(z["nb"])
    a = gen_f32(1, m * k).reshape(m, k)
    B = gen_f32(2, nb * n * k).reshape(nb, n, k)
    C0 = gen_f32(3, nb * n * m).reshape(nb, n, m)
    return m, k, n, nb, a, B, C0, float(z["thr"]), float(z["alpha"]), float(z["beta"])


def bell_case(z):
    m, k, n, nb, block, ell_cols = (int(z[x]) for x in ("m", "k", "n", "nb", "block", "ell_cols"))
    V = gen_f32(4, nb * m * ell_cols).reshape(nb, m, ell_cols)
    B = gen_f32(5, n * k).reshape(n, k)
    return m, k, n, nb, block, ell_cols, V, B


@pytest.mark.parametrize("path", CUSP_COO, ids=os.path.basename)
def test_threshold_and_coo_spmm_oracles_are_bit_exact_with_cusparse(orc, path):
    z = np.load(path)
    m, k, n, nb, a, B, C0, thr, alpha, beta = coo_case(z)
    ri, ci, va, _ = orc.threshold_to_coo(2, a, thr)
    assert ri.size == int(z["nnz"]) and sha(np.concatenate([ri, ci])) == str(z["coo_sha256"])
    want = orc.spmm_coo_batched_f64(m, k, n, nb, ri, ci, va, B, C0, alpha, beta)
    assert np.array_equal(want.astype(np.float32), z["c"])
    assert np.array_equal(want, z["c"].astype(np.float64))  # nothing was rounded on the way


@pytest.mark.parametrize("path", CUSP_BELL, ids=os.path.basename)
def test_blocked_ell_oracle_is_bit_exact_with_cusparse(orc, path):
    z = np.load(path)
    m, k, n, nb, block, ell_cols, V, B = bell_case(z)
    for b in range(nb):
        want = orc.spmm_bell_f64(m, k, n, block, ell_cols, z["col_idx"][b], V[b], B)
        assert np.array_equal(want, z["c"][b].astype(np.float64))


@pytest.mark.parametrize("path", SPARSIFY, ids=os.path.basename)
def test_oracle_matches_reference_sparsify(orc, path):
    z = np.load(path)
    m, n, blk, sf = int(z["m"]), int(z["n"]), int(z["blk"]), float(z["sf"])
    w = (1.0 + (np.arange(m * n) % 251)).astype(np.float32)
    got_w, got_mask = orc.prune_blocks_ref(w, m, n, blk // 10, blk % 10, sf)
    if "weights" in z.files:
        assert np.array_equal(got_w, z["weights"])
        assert np.array_equal(got_mask, z["mask"])
    else:
        assert sha(got_w) == str(z["weights_sha256"])
        assert sha(got_mask) == str(z["mask_sha256"])
        assert int((got_w == 0).sum()) == int(z["zeros"]) and int(got_mask.sum()) == int(z["mask_sum"])


@pytest.mark.parametrize("path", [p for p in CUSPLT if "_strip_" in p], ids=os.path.basename)
def test_oracle_strip_prune_is_bit_exact_with_cusparselt(orc, path):
    z = np.load(path)
    m, k = int(z["m"]), int(z["k"])
    a = gen(1, m * k).view(np.uint16).reshape(m, k)
    assert int(z["valid"]) == 0
    mine = orc.prune24_strip(orc.F16, a, want_mask=False)["dense"]
    assert np.array_equal(mine, z["a_pruned"])  # same survivors, same tie-break (lower index wins)
    assert orc.prune24_check(orc.F16, z["a_pruned"]) == 0


@pytest.mark.parametrize("path", [p for p in CUSPLT if "_tile_" in p], ids=os.path.basename)
def test_oracle_tile_prune_bit_exact_with_cusparselt(orc, path):
    """TILE mode (spmma.hxx:86) is closed source; the oracle's selection rule was fitted on the tile_*.npz probes
    and must reproduce the library's pruned matrix bit for bit on these independent tie-rich inputs too."""
    z = np.load(path)
    m, k = int(z["m"]), int(z["k"])
    a = gen(1, m * k).view(np.uint16).reshape(m, k)
    mine, _ = orc.prune24_tile(orc.F16, a)
    assert np.array_equal(mine, z["a_pruned"])
    assert orc.prune24_check(orc.F16, mine) == 0


def tile_fixture(path):
    """(dtype id name, tiles [T,16] uint16, library pattern [T] uint16) of a tile_*.npz fixture"""
    from make_tile_probe import all_faces
    z = np.load(path)
    dt = str(z["dtype"])
    if "regenerate" in z.files:
        kind = str(z["regenerate"])
        if kind == "subsets":
            t = np.arange(65536, dtype=np.uint32)
            w = 1.0 + ((t[:, None] >> np.arange(16)) & 1)
        else:
            fc = np.array(all_faces(), dtype=np.uint32)
            w = 1.0 + ((fc[:, 1:2] >> np.arange(16)) & 1) + ((fc[:, 0:1] >> np.arange(16)) & 1)
        tiles = w.astype(np.float16).view(np.uint16)
    else:
        tiles = z["tiles"]
    return dt, tiles, z["pattern"]


def tiles_as_matrix(tiles):
    """[T,16] -> a (4 x 4T) matrix whose t-th 4x4 tile is tiles[t]"""
    T = tiles.shape[0]
    return np.ascontiguousarray(tiles.reshape(T, 4, 4).transpose(1, 0, 2).reshape(4, 4 * T))


TILE_FIX = sorted(glob.glob(os.path.join(GOLD, "tile_*.npz")))


def test_tile_fixtures_present():
    names = {os.path.basename(p) for p in TILE_FIX}
    assert {"tile_subsets.npz", "tile_faces.npz", "tile_small.npz", "tile_wide.npz"} <= names


@pytest.mark.parametrize("path", TILE_FIX, ids=os.path.basename)
def test_oracle_tile_selection_is_cusparselts(orc, path):
    dt, tiles, pattern = tile_fixture(path)
    assert len(tiles) == len(pattern)
    a = tiles_as_matrix(tiles)
    mine, mask = orc.prune24_tile(orc.F16 if dt == "f16" else orc.BF16, a)
    T = len(tiles)
    got = (mask.reshape(4, T, 4).transpose(1, 0, 2).reshape(T, 16).astype(np.uint32) << np.arange(16)).sum(1)
    assert np.array_equal(got.astype(np.uint16), pattern), f"{int((got != pattern).sum())} of {T} tiles differ"
    keep = ((pattern[:, None] >> np.arange(16)) & 1).astype(bool)
    want = np.where(keep, tiles, 0).astype(np.uint16)
    assert np.array_equal(mine, tiles_as_matrix(want))


@pytest.mark.parametrize("path", CUSPLT, ids=os.path.basename)
def test_gemm_oracle_agrees_with_cusparselt_matmul(orc, path):
    z = np.load(path)
    m, k, n = int(z["m"]), int(z["k"]), int(z["n"])
    b = gen(2, k * n).view(np.uint16).reshape(k, n)
    want = orc.spmma_f64(orc.F16, z["a_pruned"], b)
    got = orc.to_f32(orc.F16, z["d"]).astype(np.float64)
    scale = np.maximum(np.abs(want), 1e-2 * np.abs(want).max())
    assert float(np.max(np.abs(got - want) / scale)) <= 1e-2
    # and the timed CPU port (compressed operand, fp32 accumulate) gives the library's bits up to 1 ulp
    pr = orc.prune24_strip(orc.F16, z["a_pruned"], want_mask=False)
    port = orc.to_f32(orc.F16, orc.spmma_compressed_f32(orc.F16, pr["vals"], pr["meta"], m, k, b)).astype(np.float64)
    assert float(np.max(np.abs(port - want) / scale)) <= 1e-2
