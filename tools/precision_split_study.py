"""How much accuracy do the tensor-core operand splits cost through the WHOLE path?  (development tool; CPU only)

Every conv / linear of the oracle is replaced by an emulation of a split-operand tensor-core product with fp32 accumulation:
    tf32x3 : a = hi + lo with hi = tf32(a), lo = tf32(a - hi);   a*b ~ hi*hi + hi*lo + lo*hi     (what conv_tc / head_tc do)
    bf16x3 : the same with bf16 halves (half the TMEM bytes, kind::f16 at twice the kind::tf32 rate)
    fp16x3 : the same with fp16 halves, no scaling (11-bit halves like tf32, but a 5-bit exponent: lo halves go subnormal / flush,
             and any activation above 65504 overflows)
    tf32x1 : one tf32 product (what stock PyTorch/cuDNN does for convs by default: torch.backends.cudnn.allow_tf32 = True)
    bf16x1 : one bf16 product
and the result is compared with the fp64 oracle on the golden weights and on random_state_dict(1600):
    density  max |d - ref| / max(|ref|, 1e-3 max|ref|)      bar 1e-2   (BASELINE.json north_star)
    popcount |sum - ref| / ref                               bar 1e-3
    builtup  max abs error of the sigmoid score
    head gradients of the census step, norm-wise against the loss-term scale (po.grad_parity_errors), bar 1e-3

    python tools/precision_split_study.py > profiles/r2_split_precision_study.txt
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import popcorn_oracle as po  # noqa: E402


def tf32(x):
    # cvt.rna.tf32.f32: round to nearest, ties away, keep 10 mantissa bits
    d = x.detach()
    i = d.contiguous().view(torch.int32)
    r = ((i + 0x1000) & ~0x1FFF).view(torch.float32)
    return x + (r - d)          # straight-through: the backward kernels are separate fp32 code


def bf16(x):
    d = x.detach()
    return x + (d.to(torch.bfloat16).to(torch.float32) - d)


def fp16(x):
    d = x.detach()
    return x + (d.to(torch.float16).to(torch.float32) - d)


def split(x, rnd):
    hi = rnd(x)
    return hi, rnd(x - hi)


def make(kind):
    rnd = {"tf32": tf32, "bf16": bf16, "fp16": fp16}[kind[:4]]
    passes = int(kind[-1])

    def prod(op, a, b):
        if passes == 1:
            return op(rnd(a), rnd(b))
        ah, al = split(a, rnd)
        bh, bl = split(b, rnd)
        return op(ah, bh) + (op(ah, bl) + op(al, bh))
    return prod


class Patched:
    def __init__(self, kind):
        self.prod = make(kind) if kind != "fp32" else None

    def __enter__(self):
        if self.prod is None:
            return
        self.c2, self.ct, self.li = F.conv2d, F.conv_transpose2d, F.linear
        prod, c2, ct, li = self.prod, self.c2, self.ct, self.li

        def conv2d(x, w, b=None, **kw):
            y = prod(lambda a, c: c2(a.double(), c.double(), None, **kw).float(), x, w)
            return y if b is None else y + b.view(1, -1, 1, 1)

        def convt(x, w, b=None, **kw):
            return ct(x, w, b, **kw)          # ConvT 2x2 runs as fp32 FMAs in the epilogue (EPI_CONVT)

        def linear(x, w, b=None):
            y = prod(lambda a, c: li(a.double(), c.double()).float(), x, w)
            return y if b is None else y + b
        F.conv2d, F.conv_transpose2d, F.linear = conv2d, convt, linear

    def __exit__(self, *a):
        if self.prod is not None:
            F.conv2d, F.conv_transpose2d, F.linear = self.c2, self.ct, self.li


def fold(sd):
    """The kernels multiply BN-folded weights; fold them here so the split sees the same numbers."""
    out = dict(sd)
    for k in list(sd):
        if k.endswith(".running_var"):
            bn = k[: -len(".running_var")]
            pfx, slot = bn.rsplit(".", 1)
            conv = f"{pfx}.{int(slot) - 1}"
            g = sd[bn + ".weight"].double() / torch.sqrt(sd[bn + ".running_var"].double() + po.BN_EPS)
            out[conv + ".weight"] = (sd[conv + ".weight"].double() * g.view(-1, 1, 1, 1)).float()
            out[conv + ".bias"] = ((sd[conv + ".bias"].double() - sd[bn + ".running_mean"].double()) * g + sd[bn + ".bias"].double()).float()
            out[bn + ".weight"] = torch.ones_like(sd[bn + ".weight"])
            out[bn + ".bias"] = torch.zeros_like(sd[bn + ".bias"])
            out[bn + ".running_mean"] = torch.zeros_like(sd[bn + ".running_mean"])
            out[bn + ".running_var"] = torch.ones_like(sd[bn + ".running_var"]) - po.BN_EPS
    return out


def rel(a, b, floor=1e-3):
    a, b = a.double(), b.double()
    return float(((a - b).abs() / torch.clamp(b.abs(), min=floor * float(b.abs().max()))).max())


def main():
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    g = np.load(os.path.join(ROOT, "tests/golden/state_dict.npz"))
    sets = {"golden": {k: torch.from_numpy(g[k]) for k in g.files}, "random1600": po.random_state_dict(seed=1600)}
    H, W = 192, 256
    x = po.synthetic_input(H, W, seed=1610)
    B, h, w = 2, 96, 128
    xs = po.synthetic_input(h, w, seed=3, B=B)
    admin = torch.zeros(B, h, w)
    admin[0, 10:70, 20:100] = 4.0
    admin[1, 30:90, 8:64] = 9.0
    cidx = torch.tensor([4, 9])
    y = torch.tensor([2500.0, 9000.0])
    torch.manual_seed(7)
    grid = po.sparsity_grid(h, w)
    print(f"{'weights':11s} {'scheme':7s} {'density':>9s} {'popcount':>9s} {'builtup':>9s} {'grad(norm)':>10s} {'grad(elem)':>10s}")
    for wname, sd in sets.items():
        sdf = fold(sd)
        sd64 = {k: v.double() for k, v in sdf.items()}
        ref = po.forward(sd64, {"input": x.double()}, padding=False)
        bu_ref = po.building_score(sd64, x.double())
        inp = lambda dt: {"input": xs.to(dt), "admin_mask": admin.to(dt), "census_idx": cidx}
        total, per, _ = po.head_grad_terms(sd64, inp(torch.float64), y.double(), grid=grid, padding=False)
        for kind in ("fp32", "tf32x3", "fp16x3", "bf16x3", "tf32x1", "bf16x1"):
            with Patched(kind), torch.no_grad():
                out = po.forward(sdf, {"input": x.clone()}, padding=False)
                bu = po.building_score(sdf, x.clone())
            with Patched(kind):
                got, _, _ = po.head_grad_terms(sdf, inp(torch.float32), y, grid=grid, padding=False)
            gn, ge = po.grad_parity_errors(got, total, per)
            print(f"{wname:11s} {kind:7s} {rel(out['popdensemap'], ref['popdensemap']):9.2e} "
                  f"{rel(out['popcount'], ref['popcount'], floor=1.0):9.2e} {float((bu.double() - bu_ref).abs().max()):9.2e} {gn:10.2e} {ge:10.2e}")


if __name__ == "__main__":
    main()
