"""Implicit GEMM (spfy_spmma_conv): the `unfold` that produces every CSV shape
(reference datasets/get_shapes.py:29-41) fused into the spmma B-operand load with TMA im2col.

Parity bar: the explicit path -- torch unfold of the same activations, K reordered to (kh, kw, c), then
spfy_spmma on the materialised K x N matrix -- runs the SAME MMAs on the SAME operand bytes, so the two results must be
identical bit for bit; the explicit path itself is tied to the fp64 oracle in tests/test_gpu_parity.py."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def explicit_b(x_nhwc, kh, kw, stride, pad):
    """[K = kh*kw*c, N = batch*ho*wo] row-major, K ordered (kh, kw, c), N ordered (image, row, column)"""
    nb, h, w, c = x_nhwc.shape
    cols = F.unfold(x_nhwc.permute(0, 3, 1, 2).float(), (kh, kw), padding=pad, stride=stride)  # [nb, c*kh*kw, L], (c, kh, kw)
    L = cols.shape[2]
    cols = cols.view(nb, c, kh * kw, L).permute(2, 1, 0, 3).reshape(kh * kw * c, nb * L)       # (tap, c) x (image, position)
    return cols.to(x_nhwc.dtype).contiguous()


# 3 x 3 layers of datasets/resnet18.csv (m = ho*wo, n = C_out, k = 9*C_in) and the strided / 1 x 1 variants of the other tables
CONV_CASES = [  # (batch, h, w, c_in, c_out, kh, kw, stride, pad)
    (4, 56, 56, 64, 64, 3, 3, 1, 1),     # 3136,64,576
    (4, 56, 56, 64, 128, 3, 3, 2, 1),    # 784,128,576   (stride 2)
    (4, 28, 28, 128, 128, 3, 3, 1, 1),   # 784,128,1152
    (2, 28, 28, 128, 256, 3, 3, 2, 1),   # 196,256,1152
    (2, 14, 14, 256, 256, 3, 3, 1, 1),   # 196,256,2304
    (8, 7, 7, 512, 512, 3, 3, 1, 1),     # 49,512,4608
    (2, 56, 56, 64, 128, 1, 1, 2, 0),    # 1 x 1 stride-2 downsample
    (3, 10, 12, 64, 72, 3, 3, 1, 1),     # ragged: N = 360 is not a multiple of 128, rows wrap inside a tile
    (4, 9, 9, 192, 64, 5, 3, 2, 2),      # rectangular filter, K = 15 * 192 = 45 pieces of 64 channels (odd), N = 120
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "x".join(str(v) for v in c))
@pytest.mark.parametrize("tdt", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
def test_implicit_gemm_equals_explicit_unfold(spfy, cuda, case, tdt):
    nb, h, w, c, cout, kh, kw, stride, pad = case
    gen = torch.Generator(device=cuda)
    gen.manual_seed(sum(case))
    x = (torch.rand(nb, h, w, c, device=cuda, generator=gen) * 2 - 1).to(tdt)
    wt = (torch.rand(cout, c * kh * kw, device=cuda, generator=gen) * 2 - 1).to(tdt)  # (c, kh, kw) columns like a torch conv
    wp = spfy.permute_conv_weights(wt, c, kh, kw)
    assert torch.equal(wp.view(cout, kh * kw, c), wt.view(cout, c, kh * kw).transpose(1, 2))
    comp = spfy.prune24(wp)
    b = explicit_b(x, kh, kw, stride, pad)
    want = spfy.spmma_compressed(comp, b)
    got = spfy.spmma_conv(comp, x, kh, kw, stride=stride, pad=pad)
    torch.cuda.synchronize()
    assert got.shape == want.shape
    assert torch.equal(got, want), float((got.float() - want.float()).abs().max())


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "x".join(str(v) for v in c))
def test_implicit_gemm_nhwc_output_is_the_transpose(spfy, cuda, case):
    """spfy_spmma_conv_nhwc: the same accumulators written as [batch, ho, wo, c_out] -- bit for bit the transpose of the
    row-major result, odd tile edges included (c_out = 72: the last 8 channels of a 32-row store box are clipped)"""
    nb, h, w, c, cout, kh, kw, stride, pad = case
    gen = torch.Generator(device=cuda)
    gen.manual_seed(sum(case) + 1)
    x = (torch.rand(nb, h, w, c, device=cuda, generator=gen) * 2 - 1).half()
    wt = (torch.rand(cout, c * kh * kw, device=cuda, generator=gen) * 2 - 1).half()
    comp = spfy.prune24(spfy.permute_conv_weights(wt, c, kh, kw))
    want = spfy.spmma_conv(comp, x, kh, kw, stride=stride, pad=pad)            # [cout, N]
    got = spfy.spmma_conv_nhwc(comp, x, kh, kw, stride=stride, pad=pad)       # [nb, ho, wo, cout]
    torch.cuda.synchronize()
    assert got.shape[0] == nb and got.shape[3] == cout
    assert torch.equal(got.view(-1, cout), want.t())


def test_two_convolutions_chained_in_nhwc(spfy, cuda):
    """layer 1's NHWC output is layer 2's input as it stands: no transpose, no unfold between them (checked against
    torch's conv2d applied twice to the pruned weights)"""
    nb, h, w, c0, c1, c2 = 2, 28, 28, 64, 128, 64
    gen = torch.Generator(device=cuda)
    gen.manual_seed(11)
    x = (torch.rand(nb, h, w, c0, device=cuda, generator=gen) * 2 - 1).half()
    w1 = ((torch.rand(c1, c0 * 9, device=cuda, generator=gen) * 2 - 1) * 0.1).half()
    w2 = ((torch.rand(c2, c1 * 9, device=cuda, generator=gen) * 2 - 1) * 0.1).half()
    p1, p2 = spfy.permute_conv_weights(w1, c0, 3, 3), spfy.permute_conv_weights(w2, c1, 3, 3)
    d1, d2 = torch.empty_like(p1), torch.empty_like(p2)
    k1, k2 = spfy.prune24(p1, out_dense=d1), spfy.prune24(p2, out_dense=d2)
    y1 = spfy.spmma_conv_nhwc(k1, x, 3, 3, stride=1, pad=1)                    # [nb, 28, 28, c1]
    y2 = spfy.spmma_conv_nhwc(k2, y1, 3, 3, stride=2, pad=1)                   # [nb, 14, 14, c2]
    torch.cuda.synchronize()
    f1 = d1.view(c1, 3, 3, c0).permute(0, 3, 1, 2).float()
    f2 = d2.view(c2, 3, 3, c1).permute(0, 3, 1, 2).float()
    r1 = F.conv2d(x.permute(0, 3, 1, 2).float(), f1, padding=1).half().float()  # layer 1 rounds to fp16 like ours
    r2 = F.conv2d(r1, f2, padding=1, stride=2).permute(0, 2, 3, 1)
    scale = torch.clamp(r2.abs(), min=1e-2 * float(r2.abs().max()))
    assert float(((y2.float() - r2).abs() / scale).max()) <= 1e-2


def test_implicit_gemm_matches_a_real_convolution(spfy, cuda):
    """end to end against torch's own conv2d on the pruned weights (fp32 reference, max relative error 1e-2)"""
    nb, h, w, c, cout = 4, 28, 28, 128, 256
    gen = torch.Generator(device=cuda)
    gen.manual_seed(5)
    x = (torch.rand(nb, h, w, c, device=cuda, generator=gen) * 2 - 1).half()
    wt = (torch.rand(cout, c * 9, device=cuda, generator=gen) * 2 - 1).half()
    wp = spfy.permute_conv_weights(wt, c, 3, 3)
    pruned = torch.empty_like(wp)
    comp = spfy.prune24(wp, out_dense=pruned)
    got = spfy.spmma_conv(comp, x, 3, 3, stride=1, pad=1)                      # [cout, nb*h*w]
    w4 = pruned.view(cout, 3, 3, c).permute(0, 3, 1, 2).float()               # back to [cout, c, kh, kw]
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w4, padding=1)              # [nb, cout, h, w]
    ref = ref.permute(1, 0, 2, 3).reshape(cout, -1)
    scale = torch.clamp(ref.abs(), min=1e-2 * float(ref.abs().max()))
    assert float(((got.float() - ref).abs() / scale).max()) <= 1e-2


def test_implicit_gemm_rejects_what_it_cannot_gather(spfy, cuda):
    x = torch.zeros(2, 8, 8, 3, dtype=torch.float16, device=cuda)  # the first conv layer: 3 channels
    comp = spfy.prune24(torch.zeros(64, 3 * 49 + 1, dtype=torch.float16, device=cuda))
    comp.cols = 147
    with pytest.raises(spfy.SpfyError) as e:
        spfy.spmma_conv(comp, x, 7, 7, stride=2, pad=3)
    assert e.value.code == spfy.capi.E_UNSUPPORTED


def test_plan_mixes_convolution_layers_and_matrices(spfy, cuda):
    """spfy_spmma_plan_create_conv: a table whose 3 x 3 layers are implicit GEMMs next to plain matrix problems (both
    operand orientations) in ONE plan -- every output bitwise the single call's, NHWC outputs bitwise the transpose, and the
    plan issues one launch per (class, orientation) present, whatever the mix."""
    tdt = torch.float16
    gen = torch.Generator(device=cuda)
    gen.manual_seed(77)
    problems, wants = [], []
    for nb, h, w, c, cout, kh, kw, stride, pad, nhwc in [(4, 56, 56, 64, 64, 3, 3, 1, 1, False), (4, 56, 56, 64, 128, 3, 3, 2, 1, False),
                                                          (2, 28, 28, 128, 256, 3, 3, 1, 1, True), (8, 7, 7, 512, 512, 3, 3, 1, 1, False),
                                                          (3, 10, 12, 64, 72, 3, 3, 1, 1, True), (2, 56, 56, 64, 128, 1, 1, 2, 0, False)]:
        x = (torch.rand(nb, h, w, c, device=cuda, generator=gen) * 2 - 1).to(tdt)
        wt = (torch.rand(cout, c * kh * kw, device=cuda, generator=gen) * 2 - 1).to(tdt)
        comp = spfy.prune24(spfy.permute_conv_weights(wt, c, kh, kw))
        want = spfy.spmma_conv(comp, x, kh, kw, stride=stride, pad=pad)
        ho, wo = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
        out = torch.zeros(nb, ho, wo, cout, dtype=tdt, device=cuda) if nhwc else torch.zeros_like(want)
        problems.append(dict(comp=comp, b=x, out=out, conv=(kh, kw, stride, pad), out_t=nhwc))
        wants.append(want.t().reshape(nb, ho, wo, cout) if nhwc else want)
    for i, (M, K, N) in enumerate([(64, 147, 19008), (512, 200, 1600), (256, 2304, 392), (128, 1152, 520)]):
        a = (torch.rand(M, K, device=cuda, generator=gen) * 2 - 1).to(tdt)
        op_t = i % 2 == 1
        b = (torch.rand((N, K) if op_t else (K, N), device=cuda, generator=gen) * 2 - 1).to(tdt)
        comp = spfy.prune24(a)
        op_b = spfy.OP_T if op_t else spfy.OP_N
        wants.append(spfy.spmma_compressed(comp, b, op_b=op_b))
        problems.append(dict(comp=comp, b=b, out=torch.zeros(M, N, dtype=tdt, device=cuda), op_b=op_b))
    plan = spfy.SpmmaPlan(problems)
    plan.run()
    plan.run()
    torch.cuda.synchronize()
    for q, want in zip(problems, wants):
        assert torch.equal(q["out"], want)
    assert 1 <= plan.launches <= 12
    plan.close()
    # a bad descriptor is refused at creation
    bad = dict(problems[0])
    bad["b"] = torch.zeros(4, 56, 56, 48, dtype=tdt, device=cuda)  # channels not a multiple of 64
    bad["comp"] = spfy.prune24(torch.zeros(64, 9 * 48, dtype=tdt, device=cuda))
    with pytest.raises(spfy.SpfyError):
        spfy.SpmmaPlan([bad])
