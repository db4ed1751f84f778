"""Parameter container for the DDA dual-stream UNet (no torch.nn compute modules: the forward pass is
popcorn_b200's sm_100a kernels).  It reproduces the reference's module tree *by name* so that
``state_dict()`` / ``named_parameters()`` carry exactly the 158 keys per copy, in the reference order
(model/DDA_model/utils/networks.py:72-181, 253-330), and DDA ``.pt`` checkpoints
({'step','network','optimizer'}, networks.py:22-46) load unchanged.
"""
from __future__ import annotations

import os
import warnings
from typing import Iterator, List, Tuple

import torch
import torch.nn as nn

# utils/constants.py:170-179
STAGE1_FEATS, STAGE2_FEATS = 8, 16
DDA_DIR = "model/DDA_model/checkpoints/"
DDA_NAME = f"fusionda_newAug{STAGE1_FEATS}_{STAGE2_FEATS}"
DDA_LOSS_FACTOR = 0.5
SAR_BANDS, OPTICAL_BANDS = 2, 4


def checkpoint_path(epoch: int = 30, root: str = DDA_DIR) -> str:
    """networks.py:36 — '<dir>/networks/<NAME>_checkpoint<epoch>_lossweight<f>.pt' (relative to cwd)."""
    return os.path.join(root, "networks", f"{DDA_NAME}_checkpoint{epoch}_lossweight{DDA_LOSS_FACTOR}.pt")


def _double_conv(prefix: str, cin: int, cout: int) -> List[Tuple[str, str, tuple]]:
    spec = []
    for slot, (i, o) in ((0, (cin, cout)), (3, (cout, cout))):
        spec += [(f"{prefix}.{slot}.weight", "param", (o, i, 3, 3)), (f"{prefix}.{slot}.bias", "param", (o,))]
        bn = f"{prefix}.{slot + 1}"
        spec += [(f"{bn}.weight", "param", (o,)), (f"{bn}.bias", "param", (o,)),
                 (f"{bn}.running_mean", "buffer", (o,)), (f"{bn}.running_var", "buffer", (o,)),
                 (f"{bn}.num_batches_tracked", "counter", ())]
    return spec


def _stream_spec(prefix: str, cin: int) -> List[Tuple[str, str, tuple]]:
    f1, f2 = STAGE1_FEATS, STAGE2_FEATS
    spec = _double_conv(f"{prefix}.inc.conv.conv", cin, f1)
    spec += [(f"{prefix}.outc.conv.weight", "param", (1, f1, 1, 1)), (f"{prefix}.outc.conv.bias", "param", (1,))]
    spec += _double_conv(f"{prefix}.down_seq.down1.mpconv.1.conv", f1, f2)
    spec += _double_conv(f"{prefix}.down_seq.down2.mpconv.1.conv", f2, f2)
    spec += [(f"{prefix}.up_seq.up2.up.weight", "param", (f2, f2, 2, 2)), (f"{prefix}.up_seq.up2.up.bias", "param", (f2,))]
    spec += _double_conv(f"{prefix}.up_seq.up2.conv.conv", 2 * f2, f1)
    spec += [(f"{prefix}.up_seq.up1.up.weight", "param", (f1, f1, 2, 2)), (f"{prefix}.up_seq.up1.up.bias", "param", (f1,))]
    spec += _double_conv(f"{prefix}.up_seq.up1.conv.conv", 2 * f1, f1)
    return spec


def dda_spec() -> List[Tuple[str, str, tuple]]:
    """(key, kind, shape) for one DualStreamUNet, in the reference's state_dict order."""
    f1 = STAGE1_FEATS
    spec = _stream_spec("sar_stream", SAR_BANDS)
    spec += [("sar_out_conv.conv.weight", "param", (1, f1, 1, 1)), ("sar_out_conv.conv.bias", "param", (1,))]
    spec += _stream_spec("optical_stream", OPTICAL_BANDS)
    spec += [("optical_out_conv.conv.weight", "param", (1, f1, 1, 1)), ("optical_out_conv.conv.bias", "param", (1,))]
    spec += [("fusion_out_conv.conv.weight", "param", (1, 2 * f1, 1, 1)), ("fusion_out_conv.conv.bias", "param", (1,))]
    return spec


class _Node(nn.Module):
    """Name-only container (stands where the reference has UNet / DoubleConv / Sequential / Conv2d ...)."""


def _attach(root: nn.Module, dotted: str, kind: str, value: torch.Tensor) -> None:
    *path, leaf = dotted.split(".")
    node = root
    for name in path:
        if name not in node._modules:
            node.add_module(name, _Node())
        node = node._modules[name]
    if kind == "param":
        node.register_parameter(leaf, nn.Parameter(value))
    else:
        node.register_buffer(leaf, value)


class DualStreamUNetParams(nn.Module):
    """Weights of one DDA copy.  Attribute tree: .sar_stream / .optical_stream / .{sar,optical,fusion}_out_conv."""

    def __init__(self):
        super().__init__()
        self.spec = dda_spec()
        for key, kind, shape in self.spec:
            if kind == "counter":
                val = torch.zeros((), dtype=torch.int64)
            elif key.endswith("running_var"):
                val = torch.ones(shape)
            elif kind == "buffer":
                val = torch.zeros(shape)
            elif len(shape) == 4:
                fan_in = shape[1] * shape[2] * shape[3]
                val = torch.randn(shape) * (2.0 / fan_in) ** 0.5
            elif key.split(".")[-2] in ("1", "4") and key.endswith("weight"):
                val = torch.ones(shape)
            else:
                val = torch.zeros(shape)
            _attach(self, key, kind, val)
        self.disc = None              # networks.py:44
        self.patchsize = 512          # networks.py:181 (unused attribute kept for compatibility)
        self.num_params = sum(p.numel() for p in self.parameters() if p.requires_grad)

    # --- reference API -------------------------------------------------------------------------
    def bn_parameters(self) -> Iterator[nn.Parameter]:
        for key, kind, shape in self.spec:
            if kind == "param" and len(shape) == 1 and key.split(".")[-2] in ("1", "4"):
                yield self.get_parameter(key)

    def freeze_bn_layers(self) -> None:
        """networks.py:184-189: BN layers to eval (always the case here) and their affine params frozen."""
        for p in self.bn_parameters():
            p.requires_grad = False

    def conv_weight_keys(self) -> List[str]:
        """Keys the reference re-initialises when pretrained=False: every nn.Conv2d weight (3x3 convs and the
        1x1 out convs) — ConvTranspose2d is not an nn.Conv2d instance (model/popcorn.py:59-66)."""
        return [k for k, kind, shape in self.spec if kind == "param" and len(shape) == 4 and not k.endswith(".up.weight")]

    def forward(self, *a, **k):
        raise RuntimeError("DualStreamUNetParams holds weights only; the forward pass runs in popcorn_b200's CUDA "
                           "kernels via POPCORN.forward / popcorn_b200.ops.dda_forward")


def load_checkpoint(epoch: int = 30, device="cuda", path: str | None = None, strict_file: bool = False):
    """networks.py:32-46 equivalent -> (net, None, step).  If the DDA checkpoint file is absent (e.g. a test
    box without the reference tree) the net keeps its random initialisation and a warning is raised, unless
    strict_file=True."""
    net = DualStreamUNetParams()
    step = 0
    path = path or os.environ.get("POPCORN_DDA_CHECKPOINT") or checkpoint_path(epoch)
    if os.path.isfile(path):
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        net.load_state_dict(ckpt["network"], strict=False)       # networks.py:43
        step = ckpt.get("step", 0)
    elif strict_file:
        raise FileNotFoundError(path)
    else:
        warnings.warn(f"popcorn_b200: DDA checkpoint '{path}' not found — DualStreamUNet weights are randomly "
                      "initialised until load_state_dict() is called", RuntimeWarning)
    return net.to(device), None, step
