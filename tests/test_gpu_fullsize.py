"""GPU parity tests on the configurations that are TIMED ("What's weak" 1a): the exact workload
bench.py runs -- every layer of a datasets/*.csv table at b = 32 images, one batched STRIP prune+compress launch
and one SpmmaPlan -- checked layer by layer.

  * every layer, every column: against a torch fp32 matmul of the pruned weights (the dense matrix the same prune
    kernel writes), in column slabs;  max relative error <= 1e-2 (north_star tolerance for fp16 / bf16)
  * every unique shape: against the CPU oracle (fp64 accumulation over the oracle's OWN pruning of the same
    weights) on 4096 sampled columns that include the first and the last n-tile, and the pruned weights and the
    compressed operand bit for bit
ResNet-152 at the per-GPU shard of BASELINE.json configs[4] (global batch 256 over 8 GPUs = 32 images per GPU,
N-sharded on image boundaries) runs through the same check."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL_TOL = 1e-2
SLAB = 50176  # columns per verification slab


def bits_of(t):
    return t.contiguous().view(torch.int16).cpu().numpy().view(np.uint16)


def rel_err_t(got, want):
    scale = torch.clamp(want.abs(), min=1e-2 * float(want.abs().max()))
    return float(((got - want).abs() / scale).max())


def build_like_bench(spfy, cuda, csv, batch, tdt, seed=0x5EED):
    """the resident inputs, the batched prune and the plan exactly as bench.py main() builds them"""
    gemms = [spfy.shapes.to_gemm(s, "weights", batch) for s in spfy.shapes.read_shapes(csv)]
    gen = torch.Generator(device=cuda)
    gen.manual_seed(seed)
    layers = []
    for g in gemms:
        w = (torch.rand(g.M, g.K, device=cuda, generator=gen) * 2 - 1).to(tdt)
        b = (torch.rand(g.K, g.N, device=cuda, generator=gen) * 2 - 1).to(tdt)
        d = torch.zeros(g.M, g.N, device=cuda, dtype=tdt)
        layers.append((g, w, b, d, spfy.alloc_compressed(tdt, g.M, g.K, cuda)))
    spfy.prune24_batched([l[1] for l in layers], [l[4] for l in layers])
    plan = spfy.SpmmaPlan([dict(comp=comp, b=b, out=d) for g, w, b, d, comp in layers])
    plan.run()
    torch.cuda.synchronize()
    return layers, plan


def check_table(spfy, orc, cuda, csv, batch, tdt, oracle_cols=4096):
    dcode = 0 if tdt == torch.float16 else 1
    layers, plan = build_like_bench(spfy, cuda, csv, batch, tdt)
    worst, seen = 0.0, set()
    for g, w, b, d, comp in layers:
        pruned = torch.empty_like(w)
        single = spfy.prune24(w, out_dense=pruned)  # same kernel family, one matrix: dense pruned copy + compressed operand
        assert torch.equal(single.vals, comp.vals) and torch.equal(single.meta, comp.meta)  # batched == single, bitwise
        pf = pruned.float()
        for c0 in range(0, g.N, SLAB):
            c1 = min(g.N, c0 + SLAB)
            want = pf @ b[:, c0:c1].float()
            worst = max(worst, rel_err_t(d[:, c0:c1].float(), want))
        assert worst <= REL_TOL, (g, worst)
        if g in seen:
            continue
        seen.add(g)
        # CPU oracle: its own pruning of the same weights, fp64 accumulation, sampled columns incl. both ends
        w_bits = bits_of(w)
        ref = orc.prune24_strip(dcode, w_bits, want_mask=False)
        assert np.array_equal(bits_of(pruned), ref["dense"]), g
        ov, om = orc.pack_sm100(ref["vals"], ref["meta"], g.M, g.K)
        assert np.array_equal(comp.vals.cpu().numpy(), ov) and np.array_equal(comp.meta.cpu().numpy(), om), g
        rng = np.random.default_rng(g.M * 31 + g.K)
        cols = np.unique(np.concatenate([np.arange(128), np.arange(g.N - 128, g.N),
                                         rng.integers(0, g.N, oracle_cols - 256)]))
        tcols = torch.from_numpy(cols).to(cuda)
        want = orc.spmma_f64(dcode, ref["dense"], bits_of(b[:, tcols]))
        got = d[:, tcols].float().cpu().numpy().astype(np.float64)
        scale = np.maximum(np.abs(want), 1e-2 * np.abs(want).max())
        assert float(np.max(np.abs(got - want) / scale)) <= REL_TOL, g
    plan.close()
    return worst


@pytest.mark.parametrize("tdt", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
def test_resnet50_b32_exactly_as_benchmarked(spfy, orc, cuda, tdt):
    """BASELINE.json configs[1]: all 49 layers of datasets/resnet50.csv, b = 32, N up to 401 408 -- the bench line's workload"""
    worst = check_table(spfy, orc, cuda, "resnet50.csv", 32, tdt)
    assert worst <= REL_TOL


def test_resnet152_per_gpu_shard_of_batch_256(spfy, orc, cuda):
    """BASELINE.json configs[4]: datasets/resnet152.csv (151 layers) at global batch 256 over 8 GPUs -> this GPU's
    32 images"""
    lo, hi = spfy.multigpu.shard_batch(256, 8, 0)
    assert hi - lo == 32
    worst = check_table(spfy, orc, cuda, "resnet152.csv", hi - lo, torch.float16, oracle_cols=2048)
    assert worst <= REL_TOL


def test_resnet50_b32_implicit_plan_exactly_as_benchmarked(spfy, cuda):
    """bench.py's `implicit_plan`: the resnet50 table at b = 32 as ONE plan whose 16 3 x 3 layers read NHWC activations
    through TMA im2col (spfy_spmma_plan_create_conv) next to the 33 matrix problems.  Every conv layer's output must be
    bit for bit what the plan of plain matrices computes on the unfolded operand ((kh, kw, c)-ordered K), and the
    matrix layers must be untouched by the mix."""
    import math
    import torch.nn.functional as F
    tdt, batch = torch.float16, 32
    layers, plan = build_like_bench(spfy, cuda, "resnet50.csv", batch, tdt)
    plan.close()
    gen = torch.Generator(device=cuda)
    gen.manual_seed(99)
    problems, convs = [], []
    for i, (g, w, b, d, comp) in enumerate(layers):
        hw = g.N // batch
        ho = math.isqrt(hw)
        if g.K % 9 == 0 and (g.K // 9) % 64 == 0 and ho * ho == hw:
            x = (torch.rand(batch, ho, ho, g.K // 9, device=cuda, generator=gen) * 2 - 1).to(tdt)
            out = torch.zeros_like(d)
            problems.append(dict(comp=comp, b=x, out=out, conv=(3, 3, 1, 1)))
            convs.append((i, x, out))
        else:
            problems.append(dict(comp=comp, b=b, out=torch.zeros_like(d)))
    assert len(convs) == 16
    mixed = spfy.SpmmaPlan(problems)
    mixed.run()
    torch.cuda.synchronize()
    for i, (g, w, b, d, comp) in enumerate(layers):  # the matrix problems: same bits as the all-matrix plan
        if "conv" not in problems[i]:
            assert torch.equal(problems[i]["out"], d), g
    for i, x, out in convs:
        g, _, _, _, comp = layers[i]
        ho, cin = x.shape[1], x.shape[3]
        cols = F.unfold(x.permute(0, 3, 1, 2).float(), 3, padding=1).view(batch, cin, 9, ho * ho)
        bx = cols.permute(2, 1, 0, 3).reshape(g.K, g.N).to(tdt).contiguous()
        del cols
        want = spfy.spmma_compressed(comp, bx)
        assert torch.equal(out, want), g
        del bx, want
    mixed.close()
