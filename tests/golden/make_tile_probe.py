#!/usr/bin/env python3
"""Probe inputs that reveal how cusparseLt's CUSPARSELT_PRUNE_SPMMA_TILE (what the reference requests at
include/sparsify.me/spmma.hxx:86) breaks ties between equally heavy 4x4 patterns, and how it sums.

  python tests/golden/make_tile_probe.py gen <dir>     write <dir>/<name>.in.bin   (CPU, deterministic)
  gpurun -- bash tests/golden/run_tile_probe.sh        oracle/_ref/cusparselt_ref prunefile on each
  python tests/golden/make_tile_probe.py fixtures <dir> tests/golden
                                                       distil <dir>/*.tile.bin into tests/golden/tile_*.npz

Probe `subsets`: tile t (0..65535) holds |x| = 2 on the cells of bitmask t and 1 elsewhere, so the patterns
that tie for the maximum are exactly those with the largest overlap with t -- every face of the pattern
polytope appears once, and the library's choice on each is read off the output.
"""
import os
import sys

import numpy as np

N = 1024  # 256 x 256 tiles of 4x4 = 65536 tiles


def tiles_to_matrix(t):  # t: [65536, 4, 4] -> [1024, 1024]
    return t.reshape(256, 256, 4, 4).transpose(0, 2, 1, 3).reshape(N, N)


def matrix_to_tiles(a):
    return a.reshape(256, 4, 256, 4).transpose(0, 2, 1, 3).reshape(-1, 4, 4)


def f32_to_bf16_bits(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
    return u.astype(np.uint16)


def probes():
    rng = np.random.default_rng(0x5EED)
    out = {}
    t = np.arange(65536, dtype=np.uint32)
    bits = ((t[:, None] >> np.arange(16)[None, :]) & 1).reshape(-1, 4, 4).astype(np.float32)
    sub = 1.0 + bits
    out["subsets"] = ("f16", tiles_to_matrix(sub).astype(np.float16).view(np.uint16))
    sign = rng.integers(0, 2, size=(65536, 4, 4)) * 2.0 - 1.0
    out["subsets_signed"] = ("f16", tiles_to_matrix(sub * sign).astype(np.float16).view(np.uint16))
    q = rng.integers(-64, 64, size=(N, N)).astype(np.float32) / 64.0
    out["rnd64"] = ("f16", q.astype(np.float16).view(np.uint16))
    s = rng.integers(-3, 4, size=(N, N)).astype(np.float32) / 4.0
    out["small"] = ("f16", s.astype(np.float16).view(np.uint16))
    u = rng.uniform(-1, 1, size=(N, N)).astype(np.float32)
    out["full"] = ("f16", u.astype(np.float16).view(np.uint16))
    out["full_bf16"] = ("bf16", f32_to_bf16_bits(u))
    w = (rng.standard_normal((N, N)) * np.exp2(rng.integers(-12, 12, size=(N, N)))).astype(np.float32)
    out["wide"] = ("f16", w.astype(np.float16).view(np.uint16))
    out["rnd64_bf16"] = ("bf16", f32_to_bf16_bits(q))
    # second round: every face of the pattern polytope (3 on the forced cells, 2 on the free ones, 1 outside),
    # and inputs whose fp32 pattern sums round (a few huge cells + full-mantissa small ones)
    fc = all_faces()
    w = np.ones((65536, 16), np.float32)
    for i, (o, u) in enumerate(fc):
        cells = (np.array([u], np.uint32) >> np.arange(16)) & 1
        forced = (np.array([o], np.uint32) >> np.arange(16)) & 1
        w[i] = 1 + cells + forced
    w[len(fc):] = rng.integers(1, 4, size=(65536 - len(fc), 16))
    out["faces"] = ("f16", tiles_to_matrix(w.reshape(-1, 4, 4)).astype(np.float16).view(np.uint16))
    big = rng.integers(1, 4, size=(N, N)) * 2048.0
    tiny = rng.uniform(0, 4, size=(N, N))
    sgn = rng.integers(0, 2, size=(N, N)) * 2.0 - 1.0
    w2 = (np.where(rng.random((N, N)) < 0.5, big, tiny) * sgn).astype(np.float32)
    out["wide2"] = ("f16", w2.astype(np.float16).view(np.uint16))
    out["wide2_bf16"] = ("bf16", f32_to_bf16_bits(w2))
    bits16 = (rng.integers(0, 0x7C00, size=(N, N)) | (rng.integers(0, 2, size=(N, N)) << 15)).astype(np.uint16)
    out["anybits"] = ("f16", bits16)  # every finite fp16, subnormals included
    e = rng.integers(127 - 40, 127 + 40, size=(N, N)).astype(np.uint32)
    bb = ((rng.integers(0, 2, size=(N, N)).astype(np.uint32) << 15) | (e << 7) | rng.integers(0, 128, size=(N, N)).astype(np.uint32))
    out["anybits_bf16"] = ("bf16", bb.astype(np.uint16))
    big2 = np.exp2(rng.integers(0, 14, size=(N, N))) * rng.integers(1, 3, size=(N, N))
    w3 = (np.where(rng.random((N, N)) < 0.4, big2, rng.uniform(0, 1, size=(N, N)) * np.exp2(-rng.integers(0, 10, size=(N, N)))) * sgn)
    out["wide3"] = ("f16", w3.astype(np.float32).astype(np.float16).view(np.uint16))
    spec = rng.integers(0, 65536, size=(N, N)).astype(np.uint16)  # NaN / Inf / -0 included: recorded, not gated
    out["special"] = ("f16", spec)
    return out


def all_faces():
    """every face of the polytope of 4x4 0/1 matrices with row and column sums 2, as (forced-one cells, allowed
    cells) bitmask pairs: the closure of the 90 vertices under (AND, OR)"""
    import itertools
    pr = [0x3, 0x5, 0x6, 0x9, 0xA, 0xC]
    pats = []
    for p in itertools.product(pr, repeat=4):
        if all(sum((x >> c) & 1 for x in p) == 2 for c in range(4)):
            pats.append(p[0] | p[1] << 4 | p[2] << 8 | p[3] << 12)
    faces = set((p, p) for p in pats)
    frontier = list(faces)
    while frontier:
        new = []
        for (o, u) in frontier:
            for p in pats:
                f = (o & p, u | p)
                if f not in faces:
                    faces.add(f)
                    new.append(f)
        frontier = new
    return sorted(faces)


def all_patterns():
    import itertools
    pr = [0x3, 0x5, 0x6, 0x9, 0xA, 0xC]
    return np.array([p[0] | p[1] << 4 | p[2] << 8 | p[3] << 12 for p in itertools.product(pr, repeat=4)
                     if all(sum((x >> c) & 1 for x in p) == 2 for c in range(4))], dtype=np.uint16)


def to_f64(bits, dt):
    if dt == "f16":
        return bits.view(np.float16).astype(np.float64)
    return (bits.astype(np.uint32) << 16).view(np.float32).astype(np.float64)


def fixtures(d, out):
    """Keep, per probe, the tiles that say something: every tile of the two exhaustive probes (inputs are
    regenerated, only the library's choice is stored), and for the random probes the tiles on which the library
    differs from the naive chooser (exact sums, first maximum in lexicographic pattern order) plus a plain sample."""
    pats = all_patterns()
    pm = ((pats[:, None] >> np.arange(16)) & 1).astype(np.float64)
    names = dict(l.split() for l in open(os.path.join(d, "list.txt")))
    n_faces = len(all_faces())
    for name, dt in names.items():
        if name == "special":
            continue
        a = matrix_to_tiles(np.fromfile(os.path.join(d, name + ".in.bin"), np.uint16).reshape(N, N)).reshape(-1, 16)
        o = matrix_to_tiles(np.fromfile(os.path.join(d, name + ".tile.bin"), np.uint16).reshape(N, N)).reshape(-1, 16)
        keep = (o & 0x7FFF) != 0
        assert ((o == a) | ~keep).all(), "pruned values differ from the inputs"
        mag = np.abs(to_f64(a, dt))
        zero_kept_ambiguous = (keep.sum(1) != 8)
        # where an input is +-0 the library's keep bit cannot be read off the output: complete it with the
        # best-scoring valid pattern that contains the visible keeps (only the `small` probe has zeros)
        sc = mag @ pm.T
        contains = (keep.astype(np.float64) @ pm.T) == keep.sum(1)[:, None]
        lib = pats[np.where(contains, sc, -1).argmax(1)]
        assert contains.any(1).all(), "kept entries do not fit any valid pattern"
        naive = pats[sc.argmax(1)]
        if name in ("subsets", "faces"):
            cnt = 65536 if name == "subsets" else n_faces
            np.savez_compressed(os.path.join(out, f"tile_{name}.npz"), dtype=dt, regenerate=name, pattern=lib[:cnt])
            print(name, "all", cnt, "tiles; differ from naive:", int((lib[:cnt] != naive[:cnt]).sum()))
            continue
        if name == "subsets_signed":
            continue
        pick = np.nonzero((lib != naive) & ~zero_kept_ambiguous)[0][:3072]
        pick = np.union1d(pick, np.arange(1024))
        np.savez_compressed(os.path.join(out, f"tile_{name}.npz"), dtype=dt, tiles=a[pick], pattern=lib[pick])
        print(name, len(pick), "tiles; differ from naive:", int((lib[pick] != naive[pick]).sum()))


def main():
    cmd, d = sys.argv[1], sys.argv[2]
    os.makedirs(d, exist_ok=True)
    if cmd == "fixtures":
        return fixtures(d, sys.argv[3])
    if cmd == "gen":
        with open(os.path.join(d, "list.txt"), "w") as lst:
            for name, (dt, a) in probes().items():
                a.tofile(os.path.join(d, name + ".in.bin"))
                lst.write(f"{name} {dt}\n")
        print("wrote", d)


if __name__ == "__main__":
    main()
