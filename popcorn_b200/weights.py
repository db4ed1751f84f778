"""Weight packing for the sm_100a kernels: BatchNorm(eval) folding, channel-order folding and layout.

Layout contract: include/popcorn_b200.h ("Packed weights").  BN is always in eval mode on this path
(model/popcorn.py:128, 288-289), so  W' = W * g / sqrt(var + eps),  b' = (b - mean) * g / sqrt(var + eps) + beta
(SURVEY.md Appendix A); the fold is done in float64 and rounded once to fp32.
"""
from __future__ import annotations

from typing import Dict

import torch

BN_EPS = 1e-5

# (prefix inside a stream, first slot) for the ten 3x3 convs and the two transposed convs, in pack order
_LAYERS = [
    ("conv", "inc.conv.conv", 0), ("conv", "inc.conv.conv", 3),
    ("conv", "down_seq.down1.mpconv.1.conv", 0), ("conv", "down_seq.down1.mpconv.1.conv", 3),
    ("conv", "down_seq.down2.mpconv.1.conv", 0), ("conv", "down_seq.down2.mpconv.1.conv", 3),
    ("convt", "up_seq.up2.up", None),
    ("conv", "up_seq.up2.conv.conv", 0), ("conv", "up_seq.up2.conv.conv", 3),
    ("convt", "up_seq.up1.up", None),
    ("conv", "up_seq.up1.conv.conv", 0), ("conv", "up_seq.up1.conv.conv", 3),
]


def _fold_conv(sd: Dict[str, torch.Tensor], pfx: str, slot: int) -> torch.Tensor:
    w = sd[f"{pfx}.{slot}.weight"].detach().double().cpu()          # [Cout, Cin, 3, 3]
    b = sd[f"{pfx}.{slot}.bias"].detach().double().cpu()
    g = sd[f"{pfx}.{slot + 1}.weight"].detach().double().cpu()
    beta = sd[f"{pfx}.{slot + 1}.bias"].detach().double().cpu()
    mean = sd[f"{pfx}.{slot + 1}.running_mean"].detach().double().cpu()
    var = sd[f"{pfx}.{slot + 1}.running_var"].detach().double().cpu()
    s = g / torch.sqrt(var + BN_EPS)
    wf = (w * s.view(-1, 1, 1, 1)).permute(1, 2, 3, 0).reshape(-1)   # [Cin][ky][kx][Cout]
    bf = (b - mean) * s + beta
    return torch.cat([wf, bf])


def _pack_convt(sd, pfx: str) -> torch.Tensor:
    w = sd[f"{pfx}.weight"].detach().double().cpu()                  # [Cin, Cout, 2, 2]
    b = sd[f"{pfx}.bias"].detach().double().cpu()
    return torch.cat([w.permute(0, 2, 3, 1).reshape(-1), b])          # [Cin][dy][dx][Cout]


def _pad4(t: torch.Tensor) -> torch.Tensor:
    r = (-t.numel()) % 4
    return torch.cat([t, torch.zeros(r, dtype=t.dtype)]) if r else t


def pack_dda(sd: Dict[str, torch.Tensor], copy: str, tc: bool = True) -> torch.Tensor:
    """One DualStreamUNet copy ('unetmodel' | 'building_extractor') -> flat fp32 CPU tensor.
    tc=True appends the tensor-core section (include/popcorn_b200.h "Tensor-core weight section": the 3x3 conv weights
    pre-split into hi/lo halves and pre-swizzled for tcgen05), which makes pc_dda_forward run its convs on the tensor cores."""
    flat = _pack_dda_fp32(sd, copy)
    if not tc:
        return flat
    from . import _lib
    L = _lib.lib()
    base, n = L.pc_dda_tc_pack_base(), L.pc_dda_tc_pack_floats()
    out = torch.zeros(base + n, dtype=torch.float32)
    out[:flat.numel()] = flat
    _lib.check(L.pc_dda_tc_pack(flat.data_ptr(), out[base:].data_ptr()), "pc_dda_tc_pack")
    return out


def _pack_dda_fp32(sd: Dict[str, torch.Tensor], copy: str) -> torch.Tensor:
    parts = []
    for stream in ("sar_stream", "optical_stream"):
        for kind, pfx, slot in _LAYERS:
            full = f"{copy}.{stream}.{pfx}"
            parts.append(_fold_conv(sd, full, slot) if kind == "conv" else _pack_convt(sd, full))
    for oc in ("fusion_out_conv", "sar_out_conv", "optical_out_conv"):
        parts.append(_pad4(torch.cat([sd[f"{copy}.{oc}.conv.weight"].detach().double().cpu().reshape(-1),
                                      sd[f"{copy}.{oc}.conv.bias"].detach().double().cpu().reshape(-1)])))
    return torch.cat(parts).float().contiguous()


def pack_head(sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """head.{0,2,4,6} -> W1t[k][n], b1, W2t, b2, W3t, b3, w4 (row 0 of head.6), b4 (+3 pad).  Differentiable:
    built with torch ops on the live parameters so it can be re-packed each step."""
    w = lambda i: sd[f"head.{i}.weight"].flatten(1)
    b = lambda i: sd[f"head.{i}.bias"]
    parts = []
    for i in (0, 2, 4):
        parts += [w(i).t().reshape(-1), b(i)]
    parts += [w(6)[0], b(6)[0:1], torch.zeros(3, dtype=w(6).dtype, device=w(6).device)]
    return torch.cat([p.float() for p in parts]).contiguous()


def unpack_head_grad(gpack: torch.Tensor, head_in: int) -> Dict[str, torch.Tensor]:
    """Inverse of pack_head for a gradient buffer in pack layout -> per-parameter gradients."""
    out, o = {}, 0
    dims = [(head_in, 64), (64, 64), (64, 64)]
    for i, (k, n) in zip((0, 2, 4), dims):
        out[f"head.{i}.weight"] = gpack[o:o + k * n].view(k, n).t().reshape(n, k, 1, 1).contiguous()
        o += k * n
        out[f"head.{i}.bias"] = gpack[o:o + n].clone()
        o += n
    w6 = torch.zeros(2, 64, 1, 1, dtype=gpack.dtype, device=gpack.device)
    w6[0, :, 0, 0] = gpack[o:o + 64]
    o += 64
    b6 = torch.zeros(2, dtype=gpack.dtype, device=gpack.device)
    b6[0] = gpack[o]
    out["head.6.weight"], out["head.6.bias"] = w6, b6
    return out


# ---------------------------------------------------------------------------------------------------
# tcgen05 head: weights pre-split (hi/lo halves, fp16 or TF32: tc_operand_format) and pre-swizzled into the UMMA K-major SWIZZLE_128B layout
# ---------------------------------------------------------------------------------------------------
def _split_tf32(w: torch.Tensor):
    """w = hi + lo, hi exactly representable in TF32 (low 13 mantissa bits cleared), lo = w - hi (exact in fp32)."""
    hi = (w.contiguous().view(torch.int32) & -8192).view(torch.float32)     # -8192 == 0xFFFFE000
    return hi, w - hi


_SWIZZLE_DST = {}


def _swizzle_dst(K: int, device) -> torch.Tensor:
    """Flat destination index of every element of a [64, ceil(K/32)*32] matrix in the swizzled image (cached per K and device: the
    census train step re-packs the head every iteration, the index arithmetic does not change)."""
    key = (K, str(device))
    dst = _SWIZZLE_DST.get(key)
    if dst is None:
        N = 64
        katoms = (K + 31) // 32
        n = torch.arange(N, device=device).view(N, 1)
        k = torch.arange(katoms * 32, device=device).view(1, -1)
        kk = k % 32
        pos = (((kk // 4) ^ (n % 8)) * 4 + kk % 4)
        dst = ((k // 32) * (N * 32) + n * 32 + pos).reshape(-1)
        _SWIZZLE_DST[key] = dst
    return dst


def _swizzle_k_major_128b(w: torch.Tensor) -> torch.Tensor:
    """[N=64, K] fp32 -> flat image [ceil(K/32)][64 rows][32 floats]: row n of a 32-float K-atom is 128 B, its eight
    16-byte chunks XOR-swizzled with (n % 8) (Swizzle<3,4,3>); 8-row groups are 1024 B apart (the descriptor's SBO)."""
    N, K = w.shape
    assert N == 64
    katoms = (K + 31) // 32
    if katoms * 32 != K:
        wp = torch.zeros(N, katoms * 32, dtype=w.dtype, device=w.device)
        wp[:, :K] = w
    else:
        wp = w
    out = torch.empty(katoms * N * 32, dtype=w.dtype, device=w.device)
    out[_swizzle_dst(K, w.device)] = wp.reshape(-1)          # a permutation of all elements (padding columns carry zeros)
    return out


def _split_f16(w: torch.Tensor):
    """w = hi + lo with fp16 halves (11 + 11 significand bits like the TF32 split; |w| > 65504 saturates)."""
    w = w.clamp(-65504.0, 65504.0)
    hi = w.to(torch.float16)
    return hi, (w - hi.float()).clamp(-65504.0, 65504.0).to(torch.float16)


def _swizzle_k_major_128b_f16(h: torch.Tensor, slot_floats: int) -> torch.Tensor:
    """[64, K] fp16 -> atoms of [64 rows][64 halves] (128-byte rows, 16-byte chunks = 8 halves XOR-swizzled with n % 8, one atom per 64
    K elements), returned as a float32 view padded with zeros to the matrix's slot in the image (the slots keep their TF32 sizes)."""
    N, K = h.shape
    assert N == 64 and (K + 63) // 64 * 2048 <= slot_floats
    key = ("f16", K, str(h.device))
    dst = _SWIZZLE_DST.get(key)
    if dst is None:
        n = torch.arange(N, device=h.device).view(N, 1)
        k = torch.arange(K, device=h.device).view(1, -1)
        kk = k % 64
        dst = ((k // 64) * (N * 64) + n * 64 + ((kk // 8) ^ (n % 8)) * 8 + kk % 8).reshape(-1)
        _SWIZZLE_DST[key] = dst
    out = torch.zeros(2 * slot_floats, dtype=torch.float16, device=h.device)
    out[dst] = h.reshape(-1)
    return out.view(torch.float32)


def tc_operand_format() -> int:
    """0 = TF32 halves, 1 = fp16 halves: what the loaded library's tensor-core kernels multiply (pc_tc_operand_format)."""
    from . import _lib
    return int(_lib.lib().pc_tc_operand_format())


def pack_head_tc(sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """Weight image of csrc/head_tc.cu (byte offsets OFF_* there): W1hi W1lo (8 KB each, K padded to 32), W2hi W2lo
    W3hi W3lo (16 KB each), then b1 b2 b3 w4 (64 floats each) and b4 (+3 pad).  float32 tensor of 20 740 elements.
    The matrices hold TF32 or fp16 halves, whichever the library was built for (tc_operand_format); the fp16 matrices carry the layer's
    bias as K column Kpad (16 | 64)."""
    parts = []
    f16 = tc_operand_format() == 1
    for i in (0, 2, 4):
        w = sd[f"head.{i}.weight"].detach().float().flatten(1)             # [64 out, K in]  == UMMA B operand, K-major
        if f16:
            # K padded to the 16 of a kind::f16 UMMA, then the bias as one more K column: it rides in the UMMAs against a constant
            # (1, 0, ...) k-step of the A operand (csrc/head_tc.cu BIAS_MMA)
            kpad = (w.shape[1] + 15) // 16 * 16
            wb = torch.zeros(64, kpad + 1, dtype=w.dtype, device=w.device)
            wb[:, :w.shape[1]] = w
            wb[:, kpad] = sd[f"head.{i}.bias"].detach().float()
            hi, lo = _split_f16(wb)
            slot = 2048 if i == 0 else 4096
            parts += [_swizzle_k_major_128b_f16(hi, slot), _swizzle_k_major_128b_f16(lo, slot)]
            continue
        hi, lo = _split_tf32(w)
        parts += [_swizzle_k_major_128b(hi), _swizzle_k_major_128b(lo)]
    parts += [sd[f"head.{i}.bias"].detach().float() for i in (0, 2, 4)]
    parts += [sd["head.6.weight"].detach().float().flatten(1)[0], sd["head.6.bias"].detach().float()[0:1],
              torch.zeros(3, dtype=torch.float32, device=parts[0].device)]
    out = torch.cat(parts).contiguous()
    assert out.numel() * 4 == 82960, out.numel()
    return out
