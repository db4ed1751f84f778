"""CPU study: which operand precision does the 1e-2 per-pixel bar need?  (development tool; uses the oracle as the fp32 truth)

Every conv / transposed conv / linear of the hot path gets its activations AND weights rounded to `m` explicit mantissa
bits (round-to-nearest-even on the fp32 bit pattern) before an fp32-accumulated product — the arithmetic of a tensor-core
kernel whose operands carry m bits: m=10 single-pass TF32, m=7 BF16, m=16 "bf16x2" (x = b1 + b2, four bf16 MMAs), m=21
3xTF32 (what csrc/conv_tc.cu and head_tc.cu implement).  Error metric = tests/util.py: |a-b| / max(|b|, 1e-3 max|b|).

    python tools/precision_study.py [size] [seed,seed,... | golden]          # default 384, benchmark weights seed 1600
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import popcorn_oracle as po  # noqa: E402

realF = po.F


def rnd(x: torch.Tensor, m: int) -> torch.Tensor:
    if m >= 23:
        return x
    i = x.contiguous().view(torch.int32)
    drop = 23 - m
    half = (1 << (drop - 1)) - 1 + ((i >> drop) & 1)            # round to nearest even
    return ((i + half) & ~((1 << drop) - 1)).view(torch.float32)


def trunc(x: torch.Tensor, m: int) -> torch.Tensor:
    """What the tensor core does to a raw fp32 operand: the low mantissa bits are ignored (round toward zero)."""
    return (x.contiguous().view(torch.int32) & ~((1 << (23 - m)) - 1)).view(torch.float32)


class Split3:
    """x = b1 + b2 (two bf16 pieces, 16 significant bits); product = b1*w1 + b1*w2 + b2*w1 — three bf16 MMAs at twice the
    TF32 rate = HALF the tensor time of 3xTF32; the dropped b2*w2 term is 2^-16 relative."""

    def __getattr__(self, name):
        return getattr(realF, name)

    @staticmethod
    def _op(fn, x, w, b, **k):
        x1, w1 = rnd(x, 7), rnd(w, 7)
        x2, w2 = rnd(x - x1, 7), rnd(w - w1, 7)
        return fn(x1, w1, b, **k) + fn(x1, w2, None, **k) + fn(x2, w1, None, **k)

    def conv2d(self, x, w, b=None, **k):
        return self._op(realF.conv2d, x, w, b, **k)

    def conv_transpose2d(self, x, w, b=None, **k):
        return self._op(realF.conv_transpose2d, x, w, b, **k)

    def linear(self, x, w, b=None):
        return self._op(realF.linear, x, w, b)


OPERAND_MAX = [0.0]


def fp16_pair(x: torch.Tensor, m: int = 0) -> torch.Tensor:
    """x = h1 + h2 with two fp16 pieces (22 significant bits while |x| < 65504 and x - h1 is not subnormal); all four products
    (h1 + h2)(w1 + w2) = two kind::f16 UMMAs per weight image on pair-packed operands (tools/probe/conv_pair.cu)."""
    OPERAND_MAX[0] = max(OPERAND_MAX[0], float(x.abs().max()))
    h1 = x.to(torch.float16)
    return h1.float() + (x - h1.float()).to(torch.float16).float()


def bf16_pair(x: torch.Tensor, m: int = 0) -> torch.Tensor:
    b1 = x.to(torch.bfloat16)
    return b1.float() + (x - b1.float()).to(torch.bfloat16).float()


class QF:
    """torch.nn.functional with quantised operands for the contraction ops."""

    def __init__(self, m_act, m_w, q=None):
        self.ma, self.mw = m_act, m_w
        self.q = q or rnd

    def __getattr__(self, name):
        return getattr(realF, name)

    def conv2d(self, x, w, b=None, **k):
        return realF.conv2d(self.q(x, self.ma), self.q(w, self.mw), b, **k)

    def conv_transpose2d(self, x, w, b=None, **k):
        return realF.conv_transpose2d(self.q(x, self.ma), self.q(w, self.mw), b, **k)

    def linear(self, x, w, b=None):
        return realF.linear(self.q(x, self.ma), self.q(w, self.mw), b)


def rel(a, b):
    a, b = a.double(), b.double()
    return (a - b).abs() / torch.clamp(b.abs(), min=1e-3 * float(b.abs().max()))


def golden_state_dict():
    import numpy as np
    f = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "state_dict.npz")
    return {k: torch.from_numpy(np.asarray(v)) for k, v in np.load(f).items()}


if __name__ == "__main__":
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 384
    which = sys.argv[2] if len(sys.argv) > 2 else "1600"
    torch.set_num_threads(os.cpu_count() or 1)
    if which == "golden":     # the parity tests' weights: reference init (kaiming fan_out) for unetmodel + the real DDA checkpoint
        configs = [("golden", golden_state_dict(), s_) for s_ in (1610, 3)]
    else:
        configs = [(f"seed {int(v)}", po.random_state_dict(seed=int(v)), 10 + int(v)) for v in which.split(",")]
    schemes = [("bf16 single pass, RN (7/7)", QF(7, 7)), ("tf32 single pass, raw fp32 in = truncation (10/10)", QF(10, 10, trunc)),
               ("tf32 single pass, RN (10/10)", QF(10, 10)), ("activations tf32 RN, weights fp32 (2 MMAs)", QF(10, 23)),
               ("activations fp32, weights tf32 RN (2 MMAs)", QF(23, 10)), ("12/12", QF(12, 12)), ("14/14", QF(14, 14)),
               ("2 x bf16 pieces, 3 products (1.5 tf32 slots)", Split3()), ("2 x bf16 pieces, 4 products (pair-packed)", QF(0, 0, bf16_pair)),
               ("16/16", QF(16, 16)), ("18/18", QF(18, 18)), ("19/19", QF(19, 19)),
               ("3xTF32 as implemented (21/21, 3 tf32 slots)", QF(21, 21)), ("2 x fp16 pieces, 4 products (pair-packed)", QF(0, 0, fp16_pair))]
    refs = []
    for name, sd, seed in configs:
        x = po.synthetic_input(S, S, seed=seed)
        with torch.no_grad():
            ref = po.forward(sd, {"input": x.clone()}, padding=False)
            r64 = po.forward({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}, {"input": x.double()}, padding=False)
        refs.append((sd, x, ref))
        print(f"[{name}] fp32 oracle vs fp64 oracle (the noise floor of the comparison itself): density max "
              f"{float(rel(ref['popdensemap'], r64['popdensemap']).max()):.2e}")
    print(f"size {S}x{S}, weights {which}; error = |a-b| / max(|b|, 1e-3 max|b|) over all pixels, worst case")
    print(f"| {'operands (activation / weight mantissa bits)':52s} | density max | p99.9 | p99 | popcount | bar 1e-2 / 1e-3 |")
    print("|---|---|---|---|---|---|")
    for name, qf in schemes:
        worst = [0.0, 0.0, 0.0, 0.0]
        OPERAND_MAX[0] = 0.0
        for sd, x, ref in refs:
            po.F = qf
            try:
                with torch.no_grad():
                    out = po.forward(sd, {"input": x.clone()}, padding=False)
            finally:
                po.F = realF
            r = rel(out["popdensemap"], ref["popdensemap"]).flatten()
            q = torch.quantile(r[torch.randperm(r.numel())[:1_000_000]], torch.tensor([0.999, 0.99], dtype=torch.float64))
            pc = float((out["popcount"] - ref["popcount"]).abs() / ref["popcount"].abs())
            worst = [max(a, b) for a, b in zip(worst, (float(r.max()), float(q[0]), float(q[1]), pc))]
        ok = "PASS" if worst[0] < 1e-2 and worst[3] < 1e-3 else "fail"
        margin = 1e-2 / worst[0] if worst[0] > 0 else float("inf")
        extra = f" max|operand| {OPERAND_MAX[0]:.0f}" if OPERAND_MAX[0] else ""
        print(f"| {name:52s} | {worst[0]:.2e} | {worst[1]:.2e} | {worst[2]:.2e} | {worst[3]:.1e} | {ok} (x{margin:.1f}){extra} |")
