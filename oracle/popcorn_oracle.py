"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.

A functional fp32 restatement (torch CPU ops on a plain ``state_dict``) of POPCORN's
dense-prediction hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module; the
product path (``popcorn_b200``) never does and fails loudly without its CUDA library.

Parity pin: the reference ships no tests or golden vectors.  This
restatement is pinned against the *imported, unmodified* reference model in this
container by ``oracle/make_golden.py`` (max |diff| printed there, fixtures committed to
``tests/golden/``) and re-checked by ``tests/test_oracle_vs_reference.py`` whenever
``/root/reference`` is present — model forward / backward, and bit-exactly the reference's own
``Population_Dataset`` tiling / census methods, ``run_eval.Trainer.test_target`` loop,
``apply_transformations_and_normalize`` and ``utils/losses.get_loss``.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5  # nn.BatchNorm2d default, model/DDA_model/utils/networks.py:259,263

# [R,G,B,NIR,VV,VH] -> [VV,VH,B,G,R,NIR]   (model/popcorn.py:130-134, 296-300)
FUSION_ORDER = (4, 5, 2, 1, 0, 3)


# --------------------------------------------------------------------------------------
# DDA dual-stream UNet  (model/DDA_model/utils/networks.py)
# --------------------------------------------------------------------------------------
def _conv_bn_relu(sd: Dict[str, Tensor], pfx: str, ci: int, x: Tensor) -> Tensor:
    """Conv2d(3x3,pad1) -> BatchNorm2d(eval) -> ReLU; networks.py:258-267 (slots ci, ci+1)."""
    y = F.conv2d(x, sd[f"{pfx}.{ci}.weight"], sd[f"{pfx}.{ci}.bias"], padding=1)
    b = ci + 1
    y = F.batch_norm(y, sd[f"{pfx}.{b}.running_mean"], sd[f"{pfx}.{b}.running_var"],
                     sd[f"{pfx}.{b}.weight"], sd[f"{pfx}.{b}.bias"], training=False, eps=BN_EPS)
    return F.relu(y)


def _double_conv(sd, pfx, x):
    """DoubleConv, networks.py:253-271."""
    return _conv_bn_relu(sd, pfx, 3, _conv_bn_relu(sd, pfx, 0, x))


def _up(sd, pfx, x1, x2):
    """Up.forward, networks.py:305-320: ConvT(k2,s2) -> zero-pad to skip size -> cat[skip, up] -> DoubleConv."""
    x1 = F.conv_transpose2d(x1, sd[f"{pfx}.up.weight"], sd[f"{pfx}.up.bias"], stride=2)
    dy = x2.shape[2] - x1.shape[2]
    dx = x2.shape[3] - x1.shape[3]
    x1 = F.pad(x1, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))
    return _double_conv(sd, f"{pfx}.conv.conv", torch.cat([x2, x1], dim=1))


def unet_stream(sd: Dict[str, Tensor], pfx: str, z: Tensor, encoder_no_grad: bool = False) -> Tensor:
    """UNet.forward with TOPOLOGY [8,16], enable_outc=False; networks.py:121-151, utils/constants.py:173.
    encoder_no_grad: inc / down1 / down2 run under torch.no_grad() (networks.py:124-131)."""
    with torch.set_grad_enabled(torch.is_grad_enabled() and not encoder_no_grad):
        a = _double_conv(sd, f"{pfx}.inc.conv.conv", z)
        b = _double_conv(sd, f"{pfx}.down_seq.down1.mpconv.1.conv", F.max_pool2d(a, 2))
        c = _double_conv(sd, f"{pfx}.down_seq.down2.mpconv.1.conv", F.max_pool2d(b, 2))
    u = _up(sd, f"{pfx}.up_seq.up2", c, b)
    return _up(sd, f"{pfx}.up_seq.up1", u, a)


def dual_stream_features(sd, copy: str, x_fusion: Tensor, S1=True, S2=True, encoder_no_grad: bool = False) -> Tensor:
    """DualStreamUNet.forward(..., return_features=True); networks.py:192-211."""
    feats = []
    if S1:
        feats.append(unet_stream(sd, f"{copy}.sar_stream", x_fusion[:, :2], encoder_no_grad))
    if S2:
        feats.append(unet_stream(sd, f"{copy}.optical_stream", x_fusion[:, 2:], encoder_no_grad))
    return torch.cat(feats, dim=1)


def dual_stream_logits(sd, copy: str, x_fusion: Tensor, S1=True, S2=True) -> Tensor:
    """DualStreamUNet.forward(..., return_features=False) -> the logits POPCORN uses; networks.py:213-237."""
    f = dual_stream_features(sd, copy, x_fusion, S1, S2)
    if S1 and S2:
        name = "fusion_out_conv"
    elif S1:
        name = "sar_out_conv"
    else:
        name = "optical_out_conv"
    return F.conv2d(f, sd[f"{copy}.{name}.conv.weight"], sd[f"{copy}.{name}.conv.bias"])


# --------------------------------------------------------------------------------------
# POPCORN helpers  (model/popcorn.py)
# --------------------------------------------------------------------------------------
def modality_flags(input_channels: int) -> Tuple[bool, bool]:
    """model/popcorn.py:47-54."""
    if input_channels == 0:
        return False, False
    if input_channels == 2:
        return True, False
    if input_channels == 4:
        return False, True
    return True, True


def to_fusion_order(x: Tensor, S1: bool, S2: bool) -> Tensor:
    """model/popcorn.py:129-146 / 295-314."""
    if S1 and S2:
        return torch.cat([x[:, 4:6], torch.flip(x[:, :3], dims=(1,)), x[:, 3:4]], dim=1)
    if S1:
        return torch.cat([x, torch.zeros(x.shape[0], 4, x.shape[2], x.shape[3], dtype=x.dtype)], dim=1)
    return torch.cat([torch.zeros(x.shape[0], 2, x.shape[2], x.shape[3], dtype=x.dtype),
                      torch.flip(x[:, :3], dims=(1,)), x[:, 3:4]], dim=1)


def feature_padding(H: int, W: int, force: bool, p: int = 14):
    """add_padding, model/popcorn.py:231-258.  Returns (top, bottom, left, right) reflect pads."""
    if force:
        return p, p, p, p
    top = bot = left = right = 0
    if H % 32 != 0:
        t = 64 - H % 64
        top, bot = t // 2, t - t // 2
    if W % 32 != 0:
        t = 64 - W % 64
        left, right = t // 2, t - t // 2
    return top, bot, left, right


def _reflect_pad(x, top, bot, left, right):
    # the reference pads H first, then W (popcorn.py:248-256); for reflect padding the two commute
    if top or bot:
        x = F.pad(x, (0, 0, top, bot), mode="reflect")
    if left or right:
        x = F.pad(x, (left, right, 0, 0), mode="reflect")
    return x


def building_score(sd, x: Tensor, S1=True, S2=True) -> Tensor:
    """create_building_score, model/popcorn.py:279-322: reflect 14 -> building_extractor -> sigmoid -> crop."""
    p = 14
    xp = F.pad(x, (p, p, p, p), mode="reflect")
    logits = dual_stream_logits(sd, "building_extractor", to_fusion_order(xp, S1, S2), S1, S2)
    return torch.sigmoid(logits)[:, :, p:-p, p:-p]


def unet_features(sd, x: Tensor, padding: bool, S1=True, S2=True, encoder_no_grad: bool = False) -> Tensor:
    """model/popcorn.py:126-158."""
    H, W = x.shape[2:]
    top, bot, left, right = feature_padding(H, W, force=padding)
    xq = _reflect_pad(x, top, bot, left, right)
    f = dual_stream_features(sd, "unetmodel", to_fusion_order(xq, S1, S2), S1, S2, encoder_no_grad)
    return f[:, :, top:top + H, left:left + W]


def head_mlp(sd, feats_flat: Tensor) -> Tensor:
    """self.head (4x 1x1 conv + ReLU), model/popcorn.py:79-85, on [n, C] rows; returns [n, 2]."""
    h = feats_flat
    for i in (0, 2, 4):
        h = F.relu(F.linear(h, sd[f"head.{i}.weight"].flatten(1), sd[f"head.{i}.bias"]))
    return F.linear(h, sd["head.6.weight"].flatten(1), sd["head.6.bias"])


def sparsity_grid(H: int, W: int, sub: int = 60):
    """The CPU-RNG row/col grid of get_sparsity_mask, model/popcorn.py:366-368 (same RNG calls, same order)."""
    xi = torch.ones(H).multinomial(num_samples=min(sub, H), replacement=False).sort()[0]
    yi = torch.ones(W).multinomial(num_samples=min(sub, W), replacement=False).sort()[0]
    return xi, yi


def sparsity_mask(builtup: Tensor, admin_mask: Tensor, census_idx: Tensor, occupancymodel=True,
                  grid=None) -> Tensor:
    """get_sparsity_mask live branch, model/popcorn.py:361-377."""
    region = admin_mask == census_idx.view(-1, 1, 1)
    m = (builtup[:, 0] > 0) * region if occupancymodel else region.clone()
    xi, yi = grid if grid is not None else sparsity_grid(m.shape[1], m.shape[2])
    m[:, xi.unsqueeze(1), yi] = 1
    m = m * region
    if m.sum() == 0:
        m = region
    return m


def forward(sd: Dict[str, Tensor], inputs: dict, padding: bool = True, sparse: bool = False,
            occupancymodel: bool = True, sentinelbuildings: bool = True, grid=None, encoder_no_grad: bool = False,
            unet_no_grad: bool = False) -> dict:
    """POPCORN.forward, model/popcorn.py:100-193 (eval-mode BN everywhere, :128, :288-289)."""
    x = inputs["input"]
    S1, S2 = modality_flags(x.shape[1])
    if "building_counts" not in inputs or sentinelbuildings:
        with torch.no_grad():
            inputs["building_counts"] = building_score(sd, x, S1, S2)
    builtup = inputs["building_counts"]
    with torch.set_grad_enabled(torch.is_grad_enabled() and not unet_no_grad):   # popcorn.py:147-152
        feats = unet_features(sd, x, padding, S1, S2, encoder_no_grad)
    B, C, H, W = feats.shape
    aux = {}
    if sparse:
        m = sparsity_mask(builtup, inputs["admin_mask"], inputs["census_idx"], occupancymodel, grid)
        flat = feats.permute(1, 0, 2, 3).reshape(C, -1)
        mf = m.reshape(-1)
        o = torch.zeros(2, B * H * W, dtype=feats.dtype, device=feats.device)
        o[:, mf] = head_mlp(sd, flat[:, mf].t()).t()          # sparse_module_forward, popcorn.py:195-228
        out = o.view(2, B, H, W).permute(1, 0, 2, 3)[:, 0]
        aux["mask"] = m
    else:
        out = head_mlp(sd, feats.permute(0, 2, 3, 1).reshape(-1, C)).view(B, H, W, 2)[..., 0]
    if occupancymodel:
        scale = F.relu(out)                                   # popcorn.py:170
        aux["scale"] = scale[m] if sparse else scale          # popcorn.py:172-175
        dens = scale * builtup[:, 0]                          # popcorn.py:178
    else:
        dens = F.relu(out)
        aux["scale"] = None
    if "admin_mask" in inputs:                                # popcorn.py:184-190
        this_mask = inputs["admin_mask"] == inputs["census_idx"].view(-1, 1, 1)
        popcount = (dens * this_mask).sum((1, 2))
    else:
        popcount = dens.sum((1, 2))
    return {"popcount": popcount, "popdensemap": dens, **aux}


# --------------------------------------------------------------------------------------
# Loss of the census-supervised step  (utils/losses.py:12-88, run_train.py:205-213)
# --------------------------------------------------------------------------------------
def train_loss(output: dict, y: Tensor, scale_regularization: float = 0.01, lam_weak: float = 100.0) -> Tensor:
    """log_l1_loss * 1.0 + scale_regularization * mean|scale|, times lam_weak (run_train.py:205-213)."""
    loss = F.l1_loss(torch.log(output["popcount"] + 1), torch.log(y + 1))
    if output.get("scale") is not None and scale_regularization > 0:
        loss = loss + scale_regularization * output["scale"].float().abs().mean()
    return loss * lam_weak


def train_loss_terms(output: dict, y: Tensor, scale_regularization: float = 0.01, lam_weak: float = 100.0):
    """The additive terms of ``train_loss``: one log-L1 term per region of the batch (F.l1_loss averages over the
    batch, utils/losses.py:24) and the scale regulariser; ``sum(terms) == train_loss`` (tests/test_oracle_golden.py).
    Test infrastructure: gradients of the separate terms give the natural SCALE of the total gradient.  The per-region
    terms can cancel (one region over-, one under-predicted: d log(pop)/dW is ~equal for both, the signs differ), so an
    error relative to the cancelled total measures the conditioning of the batch, not the implementation."""
    B = y.numel()
    terms = [(torch.log(output["popcount"][b] + 1) - torch.log(y[b] + 1)).abs() / B * lam_weak for b in range(B)]
    if output.get("scale") is not None and scale_regularization > 0:
        terms.append(scale_regularization * output["scale"].float().abs().mean() * lam_weak)
    return terms


def grad_terms(sd: Dict[str, Tensor], inputs: dict, y: Tensor, keys, grid=None, **fw):
    """Oracle gradients of the census train step w.r.t. the tensors named in `keys`: (total {name: grad}, [per-term {name: grad}],
    output).  One forward; the terms of train_loss_terms are back-propagated one by one (test infrastructure)."""
    names = list(keys)
    sdg = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in sd.items()}
    out = forward(sdg, inputs, sparse=True, grid=grid, **fw)
    terms = train_loss_terms(out, y)
    per = []
    for t in terms:
        g = torch.autograd.grad(t, [sdg[k] for k in names], retain_graph=True, allow_unused=True)
        per.append({k: (torch.zeros_like(sdg[k]) if gi is None else gi) for k, gi in zip(names, g)})
    total = {k: sum(p[k] for p in per) for k in names}
    return total, per, out


def head_grad_terms(sd: Dict[str, Tensor], inputs: dict, y: Tensor, grid=None, **fw):
    """grad_terms for the head parameters (the census step with unet_no_grad=True, run_train.py:201-230)."""
    return grad_terms(sd, inputs, y, [k for k in sd if k.startswith("head.")], grid=grid, **fw)


def grad_parity_errors(got: Dict[str, Tensor], total: Dict[str, Tensor], per) -> Tuple[float, float]:
    """(norm-wise, element-wise) error of a gradient set against the oracle's, each normalised by the gradient scale of
    the loss TERMS: max_k ||got_k - total_k||_F / sum_t ||per_t,k||_F  and  max_k max|got_k - total_k| / max_t max|per_t,k|.
    Why not the cancelled total: see train_loss_terms.  Why not an element-wise relative error with a small floor: one
    hidden unit of one pixel whose pre-activation sits within fp32 rounding of the ReLU knee flips its contribution
    (~1/n of the term scale) between any two fp32 implementations (tools/diag_smoke.py, profiles/r2_smoke_bisect.md)."""
    e_norm = e_elem = 0.0
    for k, g in got.items():
        d = g.detach().double().cpu() - total[k].detach().double()
        sn = sum(float(p[k].double().norm()) for p in per)
        se = max(float(p[k].double().abs().max()) for p in per)
        if sn > 0:
            e_norm = max(e_norm, float(d.norm()) / sn)
        if se > 0:
            e_elem = max(e_elem, float(d.abs().max()) / se)
    return e_norm, e_elem


# --------------------------------------------------------------------------------------
# Tiling / accumulation / census aggregation
# (data/PopulationDataset.py:294-334, 656-672, 675-729, 823-852; run_eval.py:83-154)
# --------------------------------------------------------------------------------------
def get_patch_indices(h: int, w: int, patchsize: int = 2048, overlap: int = 128) -> Tensor:
    """data/PopulationDataset.py:294-316 (without the season column)."""
    stride = patchsize - overlap * 2
    x = torch.arange(0, h - patchsize, stride, dtype=int)
    y = torch.arange(0, w - patchsize, stride, dtype=int)
    main = torch.cartesian_prod(x, y).reshape(-1, 2)
    max_x, max_y = h - patchsize, w - patchsize
    bottom = torch.stack([torch.ones(len(y), dtype=int) * max_x, y]).T
    right = torch.stack([x, torch.ones(len(x), dtype=int) * max_y]).T
    corner = torch.tensor([max_x, max_y]).unsqueeze(0)
    return torch.cat([main, bottom, right, corner])


def centre_mask(ps_x: int, ps_y: int, overlap: int) -> Tensor:
    """_create_mask, data/PopulationDataset.py:656-672."""
    m = torch.zeros(ps_x, ps_y, dtype=torch.bool)
    m[overlap:ps_x - overlap, overlap:ps_y - overlap] = True
    return m


def tiled_eval(sd_list, raster: Tensor, patchsize: int = 2048, overlap: int = 128, forward_fn=None, with_scale_std: bool = False):
    """Restated run_eval.Trainer.test_target accumulate loop, run_eval.py:83-154, for one frame.

    raster [6,H,W] normalised fp32.  Returns (mean popdensemap [H,W], std map, mean scale map, count[, scale std map]).
    ``forward_fn(member, inputs)`` defaults to this module's ``forward(sd, inputs, padding=False)``;
    make_golden.py passes the imported reference model's forward instead.
    """
    if forward_fn is None:
        forward_fn = lambda sd, inp: forward(sd, inp, padding=False)
    _, h, w = raster.shape
    out = torch.zeros(h, w)
    out_sq = torch.zeros(h, w)
    out_scale = torch.zeros(h, w)
    out_scale_sq = torch.zeros(h, w)
    count = torch.zeros(h, w, dtype=torch.int16)
    mask = centre_mask(patchsize, patchsize, overlap)
    for xl, yl in get_patch_indices(h, w, patchsize, overlap).tolist():
        tile = raster[None, :, xl:xl + patchsize, yl:yl + patchsize]
        dens = torch.zeros(patchsize, patchsize)
        dens_sq = torch.zeros(patchsize, patchsize)
        scale = torch.zeros(patchsize, patchsize)
        scale_sq = torch.zeros(patchsize, patchsize)
        for sd in sd_list:
            o = forward_fn(sd, {"input": tile})
            dens += o["popdensemap"][0]
            dens_sq += o["popdensemap"][0] ** 2
            scale += o["scale"][0]
            scale_sq += o["scale"][0] ** 2
        out[xl:xl + patchsize, yl:yl + patchsize][mask] += dens[mask]
        out_sq[xl:xl + patchsize, yl:yl + patchsize][mask] += dens_sq[mask]
        out_scale[xl:xl + patchsize, yl:yl + patchsize][mask] += scale[mask]
        out_scale_sq[xl:xl + patchsize, yl:yl + patchsize][mask] += scale_sq[mask]
        count[xl:xl + patchsize, yl:yl + patchsize][mask] += len(sd_list)
    div = count > 1                                                   # run_eval.py:140-154
    cf = count[div].to(torch.float32)
    out[div] = out[div] / cf
    out_sq[div] = torch.sqrt((out_sq[div] - out[div] ** 2 * cf) / (cf - 1))
    out_scale[div] = out_scale[div] / cf
    out_scale_sq[div] = torch.sqrt((out_scale_sq[div] - out_scale[div] ** 2 * cf) / (cf - 1))
    if with_scale_std:
        return out, out_sq, out_scale, count, out_scale_sq
    return out, out_sq, out_scale, count


def convert_popmap_to_census(pred: Tensor, boundary: Tensor, census_idx, bboxes) -> Tensor:
    """data/PopulationDataset.py:696-725: per census row, sum pred over (boundary==idx) inside its bbox.

    boundary is the float32-cast id raster (:691); bbox = (xmin, xmax, ymin, ymax) or None (skipped, -> -1).
    """
    res = -torch.ones(len(census_idx), dtype=torch.float32)
    for i, (cidx, bbox) in enumerate(zip(census_idx, bboxes)):
        if bbox is None:
            continue
        x0, x1, y0, y1 = bbox
        res[i] = pred[x0:x1, y0:y1][boundary[x0:x1, y0:y1] == cidx].to(torch.float32).sum()
    return res


def adjust_map_to_census(pred: Tensor, boundary: Tensor, census_idx, bboxes, pop) -> Tensor:
    """data/PopulationDataset.py:842-850 (dasymetric rescale, in place on a clone)."""
    pred = pred.clone()
    for i, (cidx, bbox) in enumerate(zip(census_idx, bboxes)):
        x0, x1, y0, y1 = bbox
        sel = boundary[x0:x1, y0:y1] == cidx
        s = pred[x0:x1, y0:y1][sel].to(torch.float32).sum()
        if s == 0:
            continue
        pred[x0:x1, y0:y1][sel] *= pop[i] / s
    return pred


def region_sums(pred: Tensor, ids: Tensor, R: int) -> Tensor:
    """The same quantity as convert_popmap_to_census for ids 0..R-1 in one pass (fp64 accumulate)."""
    return torch.zeros(R, dtype=torch.float64).index_add_(0, ids.reshape(-1).long(), pred.reshape(-1).double())


# --------------------------------------------------------------------------------------
# Synthetic weights / inputs shared by tests and bench
# --------------------------------------------------------------------------------------
def _stream_keys(cin: int):
    dc = lambda p, i, o: [(f"{p}.0", i, o), (f"{p}.3", o, o)]
    convs = (dc("inc.conv.conv", cin, 8) + dc("down_seq.down1.mpconv.1.conv", 8, 16)
             + dc("down_seq.down2.mpconv.1.conv", 16, 16) + dc("up_seq.up2.conv.conv", 32, 8)
             + dc("up_seq.up1.conv.conv", 16, 8))
    ups = [("up_seq.up2.up", 16), ("up_seq.up1.up", 8)]
    return convs, ups


def random_state_dict(seed: int = 1600, biasinit: float = 0.9407, head_in: int = 16) -> Dict[str, Tensor]:
    """A synthetic 324-key POPCORN state_dict (key grammar: SURVEY.md Appendix A).

    Not the reference RNG stream: kaiming-like conv weights, non-trivial BN statistics so that BN
    folding is exercised; both DDA copies get independent weights.  Used where the real DDA
    checkpoint is not available (GPU box) — 'random-init DDA + occupancy head' of BASELINE configs.
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    rn = lambda *s: torch.randn(*s, generator=g)
    for copy in ("unetmodel", "building_extractor"):
        for stream, cin in (("sar_stream", 2), ("optical_stream", 4)):
            convs, ups = _stream_keys(cin)
            p = f"{copy}.{stream}"
            for name, i, o in convs:
                sd[f"{p}.{name}.weight"] = rn(o, i, 3, 3) * math.sqrt(2.0 / (9 * i))
                sd[f"{p}.{name}.bias"] = rn(o) * 0.05
                bn = name[:-1] + str(int(name[-1]) + 1)
                sd[f"{p}.{bn}.weight"] = 1.0 + 0.1 * rn(o)
                sd[f"{p}.{bn}.bias"] = 0.1 * rn(o)
                sd[f"{p}.{bn}.running_mean"] = 0.1 * rn(o)
                sd[f"{p}.{bn}.running_var"] = 0.6 + 0.8 * torch.rand(o, generator=g)
                sd[f"{p}.{bn}.num_batches_tracked"] = torch.tensor(175440, dtype=torch.int64)
            for name, c in ups:
                sd[f"{p}.{name}.weight"] = rn(c, c, 2, 2) * math.sqrt(1.0 / c)
                sd[f"{p}.{name}.bias"] = rn(c) * 0.05
            sd[f"{p}.outc.conv.weight"] = rn(1, 8, 1, 1) * 0.3
            sd[f"{p}.outc.conv.bias"] = rn(1) * 0.1
        for oc, c in (("sar_out_conv", 8), ("optical_out_conv", 8), ("fusion_out_conv", 16)):
            sd[f"{copy}.{oc}.conv.weight"] = rn(1, c, 1, 1) * 0.3
            sd[f"{copy}.{oc}.conv.bias"] = rn(1) * 0.1
    dims = [(head_in, 64), (64, 64), (64, 64), (64, 2)]
    for li, (i, o) in zip((0, 2, 4, 6), dims):
        bound = 1.0 / math.sqrt(i)
        sd[f"head.{li}.weight"] = (torch.rand(o, i, 1, 1, generator=g) * 2 - 1) * bound
        sd[f"head.{li}.bias"] = (torch.rand(o, generator=g) * 2 - 1) * bound
    sd["head.6.bias"] = biasinit * torch.ones(2)
    return sd


# dataset_stats.json:36-81 (S2 R,G,B,NIR ; S1 VV,VH) — means / stds used to draw then normalise inputs
S2_MEAN = (1460.46, 1468.30, 1383.46, 2226.68)
S2_STD = (1130.79, 1129.03, 1053.32, 1724.32)
S1_MEAN = (-11.426, -17.753)
S1_STD = (5.598, 5.008)


def synthetic_input(H: int, W: int, seed: int = 1610, B: int = 1, coarse: int = 64) -> Tensor:
    """[B,6,H,W] normalised fp32 in reference channel order [R,G,B,NIR,VV,VH] with low-frequency structure."""
    g = torch.Generator().manual_seed(seed)
    hc, wc = (H + coarse - 1) // coarse + 1, (W + coarse - 1) // coarse + 1
    low = F.interpolate(torch.randn(B, 6, hc, wc, generator=g), size=(H, W), mode="bilinear", align_corners=True)
    raw = 0.6 * low + 0.8 * torch.randn(B, 6, H, W, generator=g)
    mean = torch.tensor(S2_MEAN + S1_MEAN).view(1, 6, 1, 1)
    std = torch.tensor(S2_STD + S1_STD).view(1, 6, 1, 1)
    val = raw * std + mean
    val[:, :4] = val[:, :4].clamp(0, 10000)
    return ((val - mean) / std).contiguous()


# data/config/dataset_stats.json ("sen2springNIR", "sen1") — the values apply_normalize uses (utils/utils.py:105-127)
REF_STATS = {
    "sen2springNIR": {"mean": (1460.4567, 1468.2986, 1383.4556, 2226.6821), "std": (1130.7949, 1129.0261, 1053.3217, 1724.3213)},
    "sen1": {"mean": (-11.426, -17.753), "std": (5.5983, 5.0076)},
}


def synthetic_raw(H: int, W: int, seed: int = 1610, coarse: int = 64):
    """Raw bands in their on-disk form: S2 uint16 [4,H,W] in FILE band order (B02,B03,B04,B08 = B,G,R,NIR) and S1
    float32 [2,H,W] dB (VV,VH) — what utils/01_download_mpc_country.py:108,135-137 writes."""
    g = torch.Generator().manual_seed(seed)
    hc, wc = (H + coarse - 1) // coarse + 1, (W + coarse - 1) // coarse + 1
    low = F.interpolate(torch.randn(1, 6, hc, wc, generator=g), size=(H, W), mode="bilinear", align_corners=True)[0]
    raw = 0.6 * low + 0.8 * torch.randn(6, H, W, generator=g)
    mean = torch.tensor(S2_MEAN + S1_MEAN).view(6, 1, 1)
    std = torch.tensor(S2_STD + S1_STD).view(6, 1, 1)
    val = raw * std + mean                                     # channels [R,G,B,NIR,VV,VH]
    s2_rgbn = val[:4].clamp(0, 10000).round().to(torch.int32)
    s2_file = s2_rgbn[[2, 1, 0, 3]].to(torch.uint16).contiguous()     # file order B,G,R,NIR
    return s2_file, val[4:6].contiguous()


def read_and_normalize(s2_file: Tensor, s1: Tensor) -> Tensor:
    """Restates the reference's ingest of one window: rasterio read in band order (3,2,1,4) + .astype(float32)
    (data/PopulationDataset.py:565-567, 594-604), apply_normalize (utils/utils.py:105-127: (x - mean) / std per channel,
    fp32) and the S2|S1 concatenation (utils/utils.py:162-171).  -> [1,6,H,W]"""
    s2 = s2_file[[2, 1, 0, 3]].to(torch.float32)               # S2_RGBNIR_channels = (3,2,1,4), 1-based
    st2, st1 = REF_STATS["sen2springNIR"], REF_STATS["sen1"]
    s2n = ((s2.permute(1, 2, 0) - torch.tensor(st2["mean"])) / torch.tensor(st2["std"])).permute(2, 0, 1)
    s1n = ((s1.to(torch.float32).permute(1, 2, 0) - torch.tensor(st1["mean"])) / torch.tensor(st1["std"])).permute(2, 0, 1)
    return torch.cat([s2n, s1n], 0)[None].contiguous()


def synthetic_regions(H: int, W: int, R: int = 400, seed: int = 7):
    """int32 id raster (0 = background frame, 1..R Voronoi cells) + census-style bbox table."""
    g = torch.Generator().manual_seed(seed)
    cy = torch.rand(R, generator=g) * H
    cx = torch.rand(R, generator=g) * W
    step = max(1, int(math.sqrt(H * W / 4e6)))            # evaluate Voronoi on a coarse grid, then upsample
    ys = torch.arange(0, H, step, dtype=torch.float32)
    xs = torch.arange(0, W, step, dtype=torch.float32)
    best = torch.full((len(ys), len(xs)), float("inf"))
    ids = torch.zeros(len(ys), len(xs), dtype=torch.int32)
    for r in range(R):
        d = (ys[:, None] - cy[r]) ** 2 + (xs[None, :] - cx[r]) ** 2
        upd = d < best
        best = torch.where(upd, d, best)
        ids[upd] = r + 1
    ids = ids.repeat_interleave(step, 0).repeat_interleave(step, 1)[:H, :W].contiguous()
    m = max(1, min(H, W) // 64)
    ids[:m] = 0
    ids[:, :m] = 0
    return ids


def region_bboxes(ids: Tensor, R: int):
    """census.csv-style bbox=[xmin,xmax,ymin,ymax] per id 1..R (utils/02_preprocess_rwa_shapefile.py:294-322)."""
    out = []
    for r in range(1, R + 1):
        nz = (ids == r).nonzero()
        if len(nz) == 0:
            out.append(None)
            continue
        out.append((int(nz[:, 0].min()), int(nz[:, 0].max()) + 1, int(nz[:, 1].min()), int(nz[:, 1].max()) + 1))
    return out
