#!/usr/bin/env python3
"""Regenerate the conv-as-GEMM shape tables (m,n,k,b) for ResNet-{18,34,50,101,152}.

The reference derives these with torchvision + `unfold` at input 32x3x224x224
(reference: datasets/get_shapes.py:16-41, column mapping at :68-74):
    m = H_out*W_out, n = C_out, k = C_in*kh*kw, b = image batch (32)
It walks only the nn.Conv2d modules that are not `downsample` ones and feeds each
conv's output straight into the next conv, so the stem max-pool is *skipped*
(layer1 therefore still sees 112x112 feature maps -> m = 12544).  torchvision is not
needed for that: the ResNet topologies are fixed, so this script re-derives the same
rows arithmetically.  `tests/test_datasets.py` pins the byte-exact output (md5).

Per-model files use CRLF (python csv.writer default, as in the reference);
`shapes.csv` is the ResNet-50 table with LF endings (reference: datasets/shapes.csv).
"""
import hashlib
import os
import sys

BATCH = 32
BLOCKS = {
    "resnet18": ("basic", [2, 2, 2, 2]),
    "resnet34": ("basic", [3, 4, 6, 3]),
    "resnet50": ("bottleneck", [3, 4, 6, 3]),
    "resnet101": ("bottleneck", [3, 4, 23, 3]),
    "resnet152": ("bottleneck", [3, 8, 36, 3]),
}


def out_sz(i, k, s, p):
    return (i + 2 * p - (k - 1) - 1) // s + 1


def resnet_rows(kind, layers):
    rows = []
    hw = 224

    def conv(cin, cout, k, s, p):
        nonlocal hw
        hw = out_sz(hw, k, s, p)
        rows.append((hw * hw, cout, cin * k * k, BATCH))

    conv(3, 64, 7, 2, 3)  # stem; the max-pool that follows is not a Conv2d -> skipped
    inplanes = 64
    for stage, nblk in enumerate(layers):
        planes = 64 << stage
        for blk in range(nblk):
            stride = 2 if (stage > 0 and blk == 0) else 1
            if kind == "basic":
                conv(inplanes, planes, 3, stride, 1)
                conv(planes, planes, 3, 1, 1)
                inplanes = planes
            else:  # torchvision "v1.5" bottleneck: stride on the 3x3
                conv(inplanes, planes, 1, 1, 0)
                conv(planes, planes, 3, stride, 1)
                conv(planes, planes * 4, 1, 1, 0)
                inplanes = planes * 4
    return rows


def render(rows, eol):
    return "".join(",".join(str(v) for v in r) + eol for r in [("m", "n", "k", "b")] + rows)


def generate():
    out = {}
    for name, (kind, layers) in BLOCKS.items():
        out[name + ".csv"] = render(resnet_rows(kind, layers), "\r\n")
    # LF endings and no trailing newline, exactly like the reference's file
    out["shapes.csv"] = render(resnet_rows(*BLOCKS["resnet50"]), "\n")[:-1]
    return out


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for fn, text in generate().items():
        with open(os.path.join(here, fn), "w", newline="") as f:
            f.write(text)
        print(fn, hashlib.md5(text.encode()).hexdigest())


if __name__ == "__main__":
    sys.exit(main())
