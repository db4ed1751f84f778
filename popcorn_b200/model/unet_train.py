"""Fine-tuning path of the DDA UNet feature extractor.

The reference back-propagates into ``unetmodel`` whenever a batch has fewer than 9 M pixels (run_train.py:191-202,
``unet_no_grad=False``; with ``encoder_no_grad=True`` the encoder runs under no_grad, networks.py:124-140).  BatchNorm
layers are frozen on every forward (``freeze_bn_layers``: eval mode, ``requires_grad=False``, networks.py:184-189), so
the trainable tensors are the Conv2d / ConvTranspose2d weights and biases; BN stays folded:

    W' = W * s,  b' = (b - mean) * s + beta,  s = gamma / sqrt(var + eps)      =>      dW = dW' * s,  db = db' * s.

The UNet is driven layer by layer through the C-ABI (pc_conv3x3_layer, pc_convt2x2_layer: the same sm_100a kernels the
inference schedule uses) so that every activation can be kept for the backward, which is hand-written as well
(csrc/unet_bwd.cu: wgrad with a deterministic two-stage reduction, ReLU / max-pool backward; dgrad = the forward conv
kernels on transposed, tap-flipped weights without the ReLU).  torch.autograd only carries the bookkeeping.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from .. import _lib, ops
from ..weights import BN_EPS

# (kind, prefix inside a stream, conv slot) in pack order — same as weights._LAYERS
_LAYERS = [
    ("conv", "inc.conv.conv", 0), ("conv", "inc.conv.conv", 3),
    ("conv", "down_seq.down1.mpconv.1.conv", 0), ("conv", "down_seq.down1.mpconv.1.conv", 3),
    ("conv", "down_seq.down2.mpconv.1.conv", 0), ("conv", "down_seq.down2.mpconv.1.conv", 3),
    ("convt", "up_seq.up2.up", None),
    ("conv", "up_seq.up2.conv.conv", 0), ("conv", "up_seq.up2.conv.conv", 3),
    ("convt", "up_seq.up1.up", None),
    ("conv", "up_seq.up1.conv.conv", 0), ("conv", "up_seq.up1.conv.conv", 3),
]
_STREAMS = ("sar_stream", "optical_stream")
_ENCODER_LAYERS = (0, 1, 2, 3, 4, 5)
IDENT = 0x03020100


def trainable_keys(S1: bool = True, S2: bool = True) -> List[str]:
    """Parameter names (inside one DualStreamUNet) that receive gradients, in the order the autograd function uses."""
    keys = []
    for s, on in zip(_STREAMS, (S1, S2)):
        if not on:
            continue
        for kind, pfx, slot in _LAYERS:
            base = f"{s}.{pfx}" + (f".{slot}" if kind == "conv" else "")
            keys += [base + ".weight", base + ".bias"]
    return keys


def _plane(C: int, H: int, W: int, dev) -> torch.Tensor:
    """[C,H,W] fp32 view whose row stride is a multiple of 4 floats (aligned vector / pooled stores)."""
    rs = (max(W, 1) + 3) // 4 * 4
    return torch.empty(C, max(H, 1), rs, dtype=torch.float32, device=dev)[:, :H, :W]


def _st():
    return torch.cuda.current_stream().cuda_stream


def _conv(a, w, cout, H, W, out, relu=True, b=None, b_off=(0, 0), pool=None, a_off=(0, 0), reflect=False, chmap=IDENT,
          cin_a=None, a_hw=None):
    """pc_conv3x3_layer on [C,h,w] views."""
    L = _lib.lib()
    cin_a = a.shape[0] if cin_a is None else cin_a
    aH, aW = (a.shape[1], a.shape[2]) if a_hw is None else a_hw
    _lib.check(L.pc_conv3x3_layer(
        a.data_ptr(), cin_a, a.stride(0), a.stride(1), aH, aW, a_off[0], a_off[1], 1 if reflect else 0, chmap,
        None if b is None else b.data_ptr(), 0 if b is None else b.shape[0], 0 if b is None else b.stride(0),
        0 if b is None else b.stride(1), 0 if b is None else b.shape[1], 0 if b is None else b.shape[2], b_off[0], b_off[1],
        w.data_ptr(), None, cout, 1 if relu else 0, H, W, out.data_ptr(), out.stride(0), out.stride(1),
        None if pool is None else pool.data_ptr(), 0 if pool is None else pool.stride(0), 0 if pool is None else pool.stride(1),
        _st()), "pc_conv3x3_layer")
    return out


def _wgrad(a, g, cout, H, W, grad_pack, accumulate, b=None, b_off=(0, 0), a_off=(0, 0), reflect=False, chmap=IDENT, cin_a=None,
           a_hw=None):
    L = _lib.lib()
    cin_a = a.shape[0] if cin_a is None else cin_a
    cin = cin_a + (0 if b is None else b.shape[0])
    aH, aW = (a.shape[1], a.shape[2]) if a_hw is None else a_hw
    need = L.pc_conv_wgrad_workspace_bytes(cin, cout, H, W)
    ws = ops._ws.get(need, g.device)
    _lib.check(L.pc_conv3x3_wgrad(
        a.data_ptr(), cin_a, a.stride(0), a.stride(1), aH, aW, a_off[0], a_off[1], 1 if reflect else 0, chmap,
        None if b is None else b.data_ptr(), 0 if b is None else b.shape[0], 0 if b is None else b.stride(0),
        0 if b is None else b.stride(1), 0 if b is None else b.shape[1], 0 if b is None else b.shape[2], b_off[0], b_off[1],
        g.data_ptr(), g.stride(0), g.stride(1), cout, H, W, grad_pack.data_ptr(), 1 if accumulate else 0, ws.data_ptr(),
        ws.numel(), _st()), "pc_conv3x3_wgrad")


def _relu_bwd(g, act, out=None, add=None):
    C, H, W = act.shape
    out = _plane(C, H, W, act.device) if out is None else out
    _lib.check(_lib.lib().pc_relu_backward(g.data_ptr(), g.stride(0), g.stride(1), act.data_ptr(), act.stride(0), act.stride(1),
                                           None if add is None else add.data_ptr(), 0 if add is None else add.stride(0),
                                           0 if add is None else add.stride(1), out.data_ptr(), out.stride(0), out.stride(1),
                                           C, H, W, _st()), "pc_relu_backward")
    return out


def _pool_relu_bwd(skip, gpool, act):
    C, H, W = act.shape
    out = _plane(C, H, W, act.device)
    _lib.check(_lib.lib().pc_maxpool2x2_relu_backward(
        None if skip is None else skip.data_ptr(), 0 if skip is None else skip.stride(0), 0 if skip is None else skip.stride(1),
        gpool.data_ptr(), gpool.stride(0), gpool.stride(1), act.data_ptr(), act.stride(0), act.stride(1), out.data_ptr(),
        out.stride(0), out.stride(1), C, H, W, _st()), "pc_maxpool2x2_relu_backward")
    return out


def _convt(x, w, out):
    C, Hl, Wl = x.shape
    _lib.check(_lib.lib().pc_convt2x2_layer(x.data_ptr(), C, x.stride(0), x.stride(1), Hl, Wl, w.data_ptr(), out.data_ptr(),
                                            out.stride(0), out.stride(1), _st()), "pc_convt2x2_layer")
    return out


def _convt_dgrad(gu, w, C, Hl, Wl):
    out = _plane(C, Hl, Wl, gu.device)
    _lib.check(_lib.lib().pc_convt2x2_dgrad(gu.data_ptr(), gu.stride(0), gu.stride(1), w.data_ptr(), C, Hl, Wl, out.data_ptr(),
                                            out.stride(0), out.stride(1), _st()), "pc_convt2x2_dgrad")
    return out


def _convt_wgrad(x, gu, grad_pack, accumulate):
    L = _lib.lib()
    C, Hl, Wl = x.shape
    ws = ops._ws.get(L.pc_convt_wgrad_workspace_bytes(C, Hl, Wl), x.device)
    _lib.check(L.pc_convt2x2_wgrad(x.data_ptr(), x.stride(0), x.stride(1), gu.data_ptr(), gu.stride(0), gu.stride(1), C, Hl, Wl,
                                   grad_pack.data_ptr(), 1 if accumulate else 0, ws.data_ptr(), ws.numel(), _st()),
               "pc_convt2x2_wgrad")


def _bn_constants(cache: dict, sd: Dict[str, torch.Tensor], full: str, slot: int):
    """(s, c) with BN(eval)(y) = y * s + c for the frozen BatchNorm that follows conv `slot` (s = gamma / sqrt(var + eps),
    c = beta - mean * s), computed once in fp64 and cached as fp32 ON THE NETWORK OBJECT (`cache`): the layers are frozen on every
    forward (freeze_bn_layers), so the constants only change when a checkpoint is loaded (the entry carries the tensors' versions)."""
    ts = [sd[f"{full}.{slot + 1}.{n}"] for n in ("weight", "bias", "running_mean", "running_var")]
    sig = tuple((t.data_ptr(), t._version) for t in ts)
    hit = cache.get((full, slot))
    if hit is None or hit[0] != sig:
        g, beta, mean, var = (t.detach().double() for t in ts)
        s = g / torch.sqrt(var + BN_EPS)
        hit = (sig, s.float(), (beta - mean * s).float())
        cache[(full, slot)] = hit
    return hit[1], hit[2]


class _StreamWeights:
    """Folded forward packs, dgrad packs and the BN scale of one stream (rebuilt on every forward: the weights move).  fp32 folding with
    cached BN constants: a handful of launches per layer instead of the ~15 of a float64 round trip (the step is host-bound)."""

    def __init__(self, sd: Dict[str, torch.Tensor], stream: str, params: Dict[str, torch.Tensor], bn_cache: dict):
        self.fwd, self.dgrad, self.scale, self.shape = {}, {}, {}, {}
        self._dgrad_packs = {}
        for li, (kind, pfx, slot) in enumerate(_LAYERS):
            full = f"{stream}.{pfx}"
            if kind == "conv":
                W = params[f"{full}.{slot}.weight"].detach()                   # [cout, cin, 3, 3]
                b = params[f"{full}.{slot}.bias"].detach()
                s, c = _bn_constants(bn_cache, sd, full, slot)
                Wf = (W * s.view(-1, 1, 1, 1)).permute(1, 2, 3, 0)             # [cin, ky, kx, cout] (view)
                self.fwd[li] = torch.cat([Wf.reshape(-1), torch.addcmul(c, b, s)])
                self.dgrad[li] = Wf.flip(1, 2).permute(3, 1, 2, 0)             # [cout, ky', kx', cin]: dgrad as a forward conv (view)
                self.scale[li] = s
                self.shape[li] = tuple(W.shape)
            else:
                T = params[f"{full}.weight"].detach()                          # [cin, cout, 2, 2]
                b = params[f"{full}.bias"].detach()
                self.fwd[li] = torch.cat([T.permute(0, 2, 3, 1).reshape(-1), b])
                self.shape[li] = tuple(T.shape)

    def dgrad_pack(self, li: int, lo: int = 0, hi: int | None = None) -> torch.Tensor:
        """dgrad weights of layer li as a forward pack [cout_fwd][9][cin_fwd[lo:hi]] + zero bias (built once per step and slice)."""
        key = (li, lo, hi)
        pack = self._dgrad_packs.get(key)
        if pack is None:
            Wd = self.dgrad[li][..., lo:hi]
            pack = torch.cat([Wd.reshape(-1), torch.zeros(Wd.shape[-1], device=Wd.device, dtype=Wd.dtype)])
            self._dgrad_packs[key] = pack
        return pack


def _chmap(C: int, s: int) -> int:
    # [R,G,B,NIR,VV,VH] -> sar (VV,VH) | optical (B,G,R,NIR)   popcorn.py:130-134
    if C == 6:
        return 0x00000504 if s == 0 else 0x03000102
    return 0x00000100 if C == 2 else 0x03000102


class UNetFeaturesFn(torch.autograd.Function):
    """feats = crop(DualStreamUNet(reflect_pad(x)))  with a hand-written backward for the conv / convT parameters."""

    @staticmethod
    def forward(ctx, x, pads, encoder_no_grad, S1, S2, net, *params):
        keys = trainable_keys(S1, S2)
        pdict = dict(zip(keys, params))
        sd = dict(net.state_dict(keep_vars=True))
        B, C, H, W = x.shape
        top, bot, left, right = pads
        Hv, Wv = H + top + bot, W + left + right
        H2, W2, H4, W4 = Hv // 2, Wv // 2, Hv // 2 // 2, Wv // 2 // 2
        dev = x.device
        x = x.float()
        if x.stride(3) != 1:
            x = x.contiguous()
        sids = [i for i, on in enumerate((S1, S2)) if on]
        feats = torch.empty(B, 8 * len(sids), H, W, dtype=torch.float32, device=dev)
        if not hasattr(net, "_bn_const_cache"):
            net._bn_const_cache = {}
        weights = {s: _StreamWeights(sd, _STREAMS[s], pdict, net._bn_const_cache) for s in sids}
        saved = []
        for bi in range(B):
            for si, s in enumerate(sids):
                wt = weights[s]
                cin0 = 2 if s == 0 else 4
                A = {}
                A[0] = _conv(x[bi], wt.fwd[0], 8, Hv, Wv, _plane(8, Hv, Wv, dev), a_off=(top, left), reflect=True,
                             chmap=_chmap(C, s), cin_a=cin0, a_hw=(H, W))
                A["p1"] = _plane(8, H2, W2, dev)
                A[1] = _conv(A[0], wt.fwd[1], 8, Hv, Wv, _plane(8, Hv, Wv, dev), pool=A["p1"])
                A[2] = _conv(A["p1"], wt.fwd[2], 16, H2, W2, _plane(16, H2, W2, dev))
                A["p2"] = _plane(16, H4, W4, dev)
                A[3] = _conv(A[2], wt.fwd[3], 16, H2, W2, _plane(16, H2, W2, dev), pool=A["p2"])
                A[4] = _conv(A["p2"], wt.fwd[4], 16, H4, W4, _plane(16, H4, W4, dev))
                A[5] = _conv(A[4], wt.fwd[5], 16, H4, W4, _plane(16, H4, W4, dev))
                A["u2"] = _convt(A[5], wt.fwd[6], _plane(16, 2 * H4, 2 * W4, dev))
                o2 = ((H2 - 2 * H4) // 2, (W2 - 2 * W4) // 2)                      # F.pad split, networks.py:309-312
                A[7] = _conv(A[3], wt.fwd[7], 8, H2, W2, _plane(8, H2, W2, dev), b=A["u2"], b_off=o2)
                A[8] = _conv(A[7], wt.fwd[8], 8, H2, W2, _plane(8, H2, W2, dev))
                A["u1"] = _convt(A[8], wt.fwd[9], _plane(8, 2 * H2, 2 * W2, dev))
                o1 = ((Hv - 2 * H2) // 2, (Wv - 2 * W2) // 2)
                A[10] = _conv(A[1], wt.fwd[10], 8, Hv, Wv, _plane(8, Hv, Wv, dev), b=A["u1"], b_off=o1)
                A[11] = _conv(A[10], wt.fwd[11], 8, Hv, Wv, _plane(8, Hv, Wv, dev))
                feats[bi, 8 * si: 8 * si + 8] = A[11][:, top: top + H, left: left + W]
                saved.append(A)
        ctx.meta = (B, C, H, W, pads, encoder_no_grad, sids, keys)
        ctx.x = x
        ctx.saved_acts = saved
        ctx.weights = weights
        return feats

    @staticmethod
    def backward(ctx, g_feats):
        B, C, H, W, pads, encoder_no_grad, sids, keys = ctx.meta
        top, bot, left, right = pads
        Hv, Wv = H + top + bot, W + left + right
        H2, W2, H4, W4 = Hv // 2, Wv // 2, Hv // 2 // 2, Wv // 2 // 2
        dev = g_feats.device
        x = ctx.x
        gpacks = {s: {li: torch.zeros_like(ctx.weights[s].fwd[li]) for li in range(12)} for s in sids}
        g_feats = g_feats.float()
        k = 0
        for bi in range(B):
            for si, s in enumerate(sids):
                A, wt, gp = ctx.saved_acts[k], ctx.weights[s], gpacks[s]
                k += 1
                acc = bi > 0                                   # gradients accumulate over the images of the batch
                o1 = ((Hv - 2 * H2) // 2, (Wv - 2 * W2) // 2)
                o2 = ((H2 - 2 * H4) // 2, (W2 - 2 * W4) // 2)
                g11 = torch.zeros(8, Hv, Wv, dtype=torch.float32, device=dev)      # crop backward = zero padding
                g11[:, top: top + H, left: left + W] = g_feats[bi, 8 * si: 8 * si + 8]
                q11 = _relu_bwd(g11, A[11])
                _wgrad(A[10], q11, 8, Hv, Wv, gp[11], acc)
                q10 = _relu_bwd(_conv(q11, wt.dgrad_pack(11), 8, Hv, Wv, _plane(8, Hv, Wv, dev), relu=False), A[10])
                _wgrad(A[1], q10, 8, Hv, Wv, gp[10], acc, b=A["u1"], b_off=o1)
                gcat = _conv(q10, wt.dgrad_pack(10), 16, Hv, Wv, _plane(16, Hv, Wv, dev), relu=False)
                g_a1_skip = gcat[0:8]
                g_u1 = gcat[8:16, o1[0]: o1[0] + 2 * H2, o1[1]: o1[1] + 2 * W2]
                _convt_wgrad(A[8], g_u1, gp[9], acc)
                q8 = _relu_bwd(_convt_dgrad(g_u1, wt.fwd[9], 8, H2, W2), A[8])
                _wgrad(A[7], q8, 8, H2, W2, gp[8], acc)
                q7 = _relu_bwd(_conv(q8, wt.dgrad_pack(8), 8, H2, W2, _plane(8, H2, W2, dev), relu=False), A[7])
                _wgrad(A[3], q7, 8, H2, W2, gp[7], acc, b=A["u2"], b_off=o2)
                g_u2pad = _conv(q7, wt.dgrad_pack(7, 16, 32), 16, H2, W2, _plane(16, H2, W2, dev), relu=False)
                g_u2 = g_u2pad[:, o2[0]: o2[0] + 2 * H4, o2[1]: o2[1] + 2 * W4]
                _convt_wgrad(A[5], g_u2, gp[6], acc)
                if encoder_no_grad:
                    continue                                   # inc / down1 / down2 ran under no_grad (networks.py:124-131)
                g_a3_skip = _conv(q7, wt.dgrad_pack(7, 0, 16), 16, H2, W2, _plane(16, H2, W2, dev), relu=False)
                q5 = _relu_bwd(_convt_dgrad(g_u2, wt.fwd[6], 16, H4, W4), A[5])
                _wgrad(A[4], q5, 16, H4, W4, gp[5], acc)
                q4 = _relu_bwd(_conv(q5, wt.dgrad_pack(5), 16, H4, W4, _plane(16, H4, W4, dev), relu=False), A[4])
                _wgrad(A["p2"], q4, 16, H4, W4, gp[4], acc)
                g_p2 = _conv(q4, wt.dgrad_pack(4), 16, H4, W4, _plane(16, H4, W4, dev), relu=False)
                q3 = _pool_relu_bwd(g_a3_skip, g_p2, A[3])
                _wgrad(A[2], q3, 16, H2, W2, gp[3], acc)
                q2 = _relu_bwd(_conv(q3, wt.dgrad_pack(3), 16, H2, W2, _plane(16, H2, W2, dev), relu=False), A[2])
                _wgrad(A["p1"], q2, 16, H2, W2, gp[2], acc)
                g_p1 = _conv(q2[0:8], wt.dgrad_pack(2), 8, H2, W2, _plane(8, H2, W2, dev), relu=False, b=q2[8:16])
                q1 = _pool_relu_bwd(g_a1_skip, g_p1, A[1])
                _wgrad(A[0], q1, 8, Hv, Wv, gp[1], acc)
                q0 = _relu_bwd(_conv(q1, wt.dgrad_pack(1), 8, Hv, Wv, _plane(8, Hv, Wv, dev), relu=False), A[0])
                cin0 = 2 if s == 0 else 4
                _wgrad(x[bi], q0, 8, Hv, Wv, gp[0], acc, a_off=(top, left), reflect=True, chmap=_chmap(C, s), cin_a=cin0,
                       a_hw=(H, W))
        # ---- unfold: packed gradients of the folded layers -> gradients of the nn.Parameters ----
        grads = {}
        for s in sids:
            wt, gp = ctx.weights[s], gpacks[s]
            for li, (kind, pfx, slot) in enumerate(_LAYERS):
                full = f"{_STREAMS[s]}.{pfx}"
                frozen = encoder_no_grad and li in _ENCODER_LAYERS
                if kind == "conv":
                    cout, cin = wt.shape[li][0], wt.shape[li][1]
                    if frozen:
                        grads[f"{full}.{slot}.weight"] = None
                        grads[f"{full}.{slot}.bias"] = None
                        continue
                    sc = wt.scale[li]
                    dWf = gp[li][: cin * 9 * cout].view(cin, 3, 3, cout).permute(3, 0, 1, 2)
                    grads[f"{full}.{slot}.weight"] = (dWf * sc.view(-1, 1, 1, 1)).contiguous()
                    grads[f"{full}.{slot}.bias"] = gp[li][cin * 9 * cout:] * sc
                else:
                    cin, cout = wt.shape[li][0], wt.shape[li][1]
                    grads[f"{full}.weight"] = gp[li][: cin * 4 * cout].view(cin, 2, 2, cout).permute(0, 3, 1, 2).contiguous()
                    grads[f"{full}.bias"] = gp[li][cin * 4 * cout:].clone()
        ctx.saved_acts = None
        return (None, None, None, None, None, None) + tuple(grads[key] for key in keys)


def unet_features(net, x: torch.Tensor, pads: Tuple[int, int, int, int], encoder_no_grad: bool, S1: bool, S2: bool) -> torch.Tensor:
    """Autograd-tracked features of ``net`` (a DualStreamUNetParams) for the fine-tuning path."""
    keys = trainable_keys(S1, S2)
    params = [net.get_parameter(k) for k in keys]
    return UNetFeaturesFn.apply(x, tuple(pads), bool(encoder_no_grad), bool(S1), bool(S2), net, *params)
