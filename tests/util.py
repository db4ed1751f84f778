"""Shared helpers for the test-suite (the oracle is only ever used here as the checker)."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# tolerances of BASELINE.json's north_star
TOL_PIXEL = 1e-2     # per-pixel density, relative
TOL_REGION = 1e-3    # region counts, relative
TOL_GRAD = 1e-3      # head / UNet gradients, relative to the gradient scale of the loss terms (oracle.grad_parity_errors)


def golden(name):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in np.load(os.path.join(GOLD, name + ".npz")).items()}


def golden_state_dict(device="cpu"):
    return {k: v.to(device) for k, v in golden("state_dict").items()}


def rel_err(a: torch.Tensor, b: torch.Tensor, floor_frac: float = 1e-3) -> torch.Tensor:
    """|a-b| / max(|b|, floor_frac * max|b|)  — the per-pixel relative error of SURVEY.md §7."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    floor = floor_frac * float(b.abs().max()) if b.numel() else 0.0
    return (a - b).abs() / torch.clamp(b.abs(), min=max(floor, 1e-30))


def max_rel(a, b, floor_frac=1e-3) -> float:
    r = rel_err(a, b, floor_frac)
    return float(r.max()) if r.numel() else 0.0


def build_model(sd, device="cuda", **kw):
    import warnings
    import popcorn_b200 as pb
    kwargs = dict(input_channels=6, occupancymodel=True, pretrained=False, biasinit=0.9407, sentinelbuildings=True)
    kwargs.update(kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = pb.POPCORN(device=device, **kwargs)
    m.load_state_dict(sd, strict=True)
    return m.to(device)
