"""Synthetic benchmark workload pieces that belong to the product side (bench.py, tools/): random-init weights in the
reference's checkpoint grammar and the census training loss.  There is no network for real checkpoints or rasters
(BASELINE.json: "random-init DDA + occupancy head"), so the weights are drawn here; the CPU oracle keeps its own copy of
the same generator (oracle/popcorn_oracle.py::random_state_dict) and tests/test_host_api.py checks that both produce the
same tensors — the product never imports the oracle.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from .model.dda import STAGE1_FEATS as F1, STAGE2_FEATS as F2

_COPIES = ("unetmodel", "building_extractor")
_STREAMS = (("sar_stream", 2), ("optical_stream", 4))


def _double_convs(cin: int):
    """(prefix, Cin, Cout) of the ten 3x3 convs of one stream in forward order (networks.py:121-151, 253-330)."""
    blocks = (("inc.conv.conv", cin, F1), ("down_seq.down1.mpconv.1.conv", F1, F2), ("down_seq.down2.mpconv.1.conv", F2, F2),
              ("up_seq.up2.conv.conv", 2 * F2, F1), ("up_seq.up1.conv.conv", 2 * F1, F1))
    for prefix, i, o in blocks:
        yield f"{prefix}.0", i, o
        yield f"{prefix}.3", o, o


def random_state_dict(seed: int = 1600, biasinit: float = 0.9407, head_in: int = 2 * F1) -> Dict[str, torch.Tensor]:
    """A 324-key POPCORN state_dict with He-scaled conv weights and non-trivial BatchNorm
    statistics, so that BN folding, both DDA copies and the head all do real work.  One CPU generator, fixed draw order."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    sd: Dict[str, torch.Tensor] = {}
    for copy in _COPIES:
        for stream, cin in _STREAMS:
            p = f"{copy}.{stream}"
            for name, i, o in _double_convs(cin):
                bn = f"{name[:-1]}{int(name[-1]) + 1}"
                sd[f"{p}.{name}.weight"] = rn(o, i, 3, 3) * math.sqrt(2.0 / (9 * i))
                sd[f"{p}.{name}.bias"] = rn(o) * 0.05
                sd[f"{p}.{bn}.weight"] = 1.0 + 0.1 * rn(o)
                sd[f"{p}.{bn}.bias"] = 0.1 * rn(o)
                sd[f"{p}.{bn}.running_mean"] = 0.1 * rn(o)
                sd[f"{p}.{bn}.running_var"] = 0.6 + 0.8 * torch.rand(o, generator=g)
                sd[f"{p}.{bn}.num_batches_tracked"] = torch.tensor(175440, dtype=torch.int64)
            for name, c in (("up_seq.up2.up", F2), ("up_seq.up1.up", F1)):
                sd[f"{p}.{name}.weight"] = rn(c, c, 2, 2) * math.sqrt(1.0 / c)
                sd[f"{p}.{name}.bias"] = rn(c) * 0.05
            sd[f"{p}.outc.conv.weight"] = rn(1, F1, 1, 1) * 0.3
            sd[f"{p}.outc.conv.bias"] = rn(1) * 0.1
        for oc, c in (("sar_out_conv", F1), ("optical_out_conv", F1), ("fusion_out_conv", 2 * F1)):
            sd[f"{copy}.{oc}.conv.weight"] = rn(1, c, 1, 1) * 0.3
            sd[f"{copy}.{oc}.conv.bias"] = rn(1) * 0.1
    for li, (i, o) in zip((0, 2, 4, 6), ((head_in, 64), (64, 64), (64, 64), (64, 2))):
        bound = 1.0 / math.sqrt(i)                     # nn.Conv2d default init range
        sd[f"head.{li}.weight"] = (torch.rand(o, i, 1, 1, generator=g) * 2 - 1) * bound
        sd[f"head.{li}.bias"] = (torch.rand(o, generator=g) * 2 - 1) * bound
    sd["head.6.bias"] = biasinit * torch.ones(2)       # model/popcorn.py:88
    return sd


def census_loss(output: dict, y: torch.Tensor, scale_regularization: float = 0.01, lam_weak: float = 100.0) -> torch.Tensor:
    """The census-supervised objective of run_train.py:205-213 with the README's training flags: log-L1 between predicted
    and census counts (utils/losses.py: loss=["log_l1_loss"], lam=[1.0]) + scale_regularization * mean|scale|, times
    lam_weak.  Plain torch: losses are outside the hot path (they only produce d popcount / d scale)."""
    loss = torch.nn.functional.l1_loss(torch.log(output["popcount"] + 1), torch.log(y + 1))
    if output.get("scale") is not None and scale_regularization > 0:
        loss = loss + scale_regularization * output["scale"].float().abs().mean()
    return loss * lam_weak
