"""POPCORN with the reference's model API (model/popcorn.py:13-377) and a B200-native body.

Same constructor kwargs, same ``forward(inputs, train, padding, return_features, encoder_no_grad,
unet_no_grad, sparse)`` signature, same output dict keys (``popcount``, ``popdensemap``, ``scale``),
same side effect (``inputs["building_counts"]``), same 324-key ``state_dict``.  Every tensor op of the
reference's forward is replaced by the sm_100a kernels behind include/popcorn_b200.h; PyTorch only owns
memory, streams and autograd bookkeeping.  There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch
import torch.nn as nn

from .. import ops, weights
from . import unet_train
from .dda import STAGE1_FEATS, load_checkpoint

# forward head on tcgen05 (split operands, csrc/head_tc.cu) unless POPCORN_HEAD_TC=0 selects the fp32 SIMT kernel (csrc/head.cu)
USE_TENSOR_CORE_HEAD = os.environ.get("POPCORN_HEAD_TC", "1") != "0"
_CHECK_IDS = os.environ.get("POPCORN_CHECK_IDS", "0") == "1"
FUSED_EVAL = os.environ.get("POPCORN_FUSED_EVAL", "1") != "0"      # eval forward through pc_infer_tile_fused (one library call)


class _SparseHeadFn(torch.autograd.Function):
    """Fused sparse head forward (gather -> MLP -> ReLU -> x builtup -> scatter -> popcount) with the
    hand-written backward for the head parameters (model/popcorn.py:162-187 under unet_no_grad=True)."""

    @staticmethod
    def forward(ctx, feats, builtup, idx, n_dev, n, head_in, hpack, tcpack, *params):
        # hpack / tcpack: the packed head weights (weights.pack_head / pack_head_tc of `params`), built by the caller BEFORE it waits for
        # the selected-pixel count, so that their small launches overlap the DDA passes instead of following the host sync
        if tcpack is not None:
            dens, scale_sel, pop = ops.head_sparse_forward(tcpack, feats, builtup, idx, n_dev, n, tc=True)
        else:
            dens, scale_sel, pop = ops.head_sparse_forward(hpack, feats, builtup, idx, n_dev, n)
        ctx.save_for_backward(hpack, feats, builtup if builtup is not None else torch.empty(0, device=feats.device),
                              idx, n_dev)
        ctx.n, ctx.head_in, ctx.has_bu = n, head_in, builtup is not None
        ctx.set_materialize_grads(False)
        scale_sel = scale_sel[:n]
        return pop.float(), dens, scale_sel

    @staticmethod
    def backward(ctx, g_pop, g_dens, g_scale):
        hpack, feats, builtup, idx, n_dev = ctx.saved_tensors
        builtup = builtup if ctx.has_bu else None
        B = feats.shape[0]
        if g_pop is None:
            g_pop = torch.zeros(B, dtype=torch.float32, device=feats.device)
        g_sel = None
        if g_scale is not None:
            g_sel = g_scale.float().contiguous()
        if g_dens is not None:
            # dL/dscale_i += g_dens[p_i] * builtup[p_i]   (popdensemap = scale * builtup, popcorn.py:178)
            sel = idx[:ctx.n].long()
            gd = g_dens.reshape(-1)[sel]
            if builtup is not None:
                gd = gd * builtup.reshape(-1)[sel]
            g_sel = gd if g_sel is None else g_sel + gd
        g_feats = None
        if ctx.needs_input_grad[0]:      # unetmodel is being fine-tuned: dL/dfeats at the selected pixels
            gpack, g_feats = ops.head_sparse_backward(hpack, feats, builtup, idx, n_dev, ctx.n, g_pop, 0.0, g_sel, want_g_feats=True)
        else:
            gpack = ops.head_sparse_backward(hpack, feats, builtup, idx, n_dev, ctx.n, g_pop, 0.0, g_sel)
        g = weights.unpack_head_grad(gpack, ctx.head_in)
        grads = tuple(g[f"head.{i}.{t}"] for i in (0, 2, 4, 6) for t in ("weight", "bias"))
        return (g_feats, None, None, None, None, None, None, None) + grads


class POPCORN(nn.Module):
    """POPCORN model (building extractor + occupancy head); see module docstring."""

    def __init__(self, input_channels, feature_extractor="DDA", occupancymodel=False, pretrained=False,
                 biasinit=0.75, sentinelbuildings=False, dda_checkpoint: Optional[str] = None, device=None):
        super().__init__()
        self.occupancymodel = occupancymodel
        self.sentinelbuildings = sentinelbuildings
        self.feature_extractor = feature_extractor

        self.p = 14                                    # reflect padding of the builtup pass (popcorn.py:44-45)
        self.p2d = (self.p, self.p, self.p, self.p)
        self.parent = None

        self.S1, self.S2 = True, True                  # popcorn.py:47-54
        if input_channels == 0:
            self.S1, self.S2 = False, False
        elif input_channels == 2:
            self.S1, self.S2 = True, False
        elif input_channels == 4:
            self.S1, self.S2 = False, True
        if not (self.S1 or self.S2):
            raise ValueError("POPCORN needs Sentinel-1 and/or Sentinel-2 input channels")

        if device is None:
            device = "cuda" if torch.cuda.is_available() else "cpu"   # the reference hard-codes "cuda" (popcorn.py:57)

        self.unetmodel, _, _ = load_checkpoint(epoch=30, device=device, path=dda_checkpoint)
        if not pretrained:                             # popcorn.py:59-66
            with torch.no_grad():
                for key in self.unetmodel.conv_weight_keys():
                    nn.init.kaiming_normal_(self.unetmodel.get_parameter(key), mode="fan_out", nonlinearity="relu")
                for key, kind, shape in self.unetmodel.spec:
                    if kind == "param" and len(shape) == 1 and key.split(".")[-2] in ("1", "4"):
                        nn.init.constant_(self.unetmodel.get_parameter(key), 1.0 if key.endswith("weight") else 0.0)

        head_input_dim = self.S1 * STAGE1_FEATS + self.S2 * STAGE1_FEATS
        self.head_input_dim = head_input_dim
        self.unetmodel.num_params = sum(p.numel() for p in self.unetmodel.parameters() if p.requires_grad)

        h = 64                                         # popcorn.py:79-88 (parameter holders; compute is fused CUDA)
        self.head = nn.Sequential(
            nn.Conv2d(head_input_dim, h, kernel_size=1, padding=0), nn.ReLU(inplace=True),
            nn.Conv2d(h, h, kernel_size=1, padding=0), nn.ReLU(inplace=True),
            nn.Conv2d(h, h, kernel_size=1, padding=0), nn.ReLU(inplace=True),
            nn.Conv2d(h, 2, kernel_size=1, padding=0)).to(device)
        self.head[-1].bias.data = biasinit * torch.ones(2, device=device)

        self.num_params = sum(p.numel() for p in self.head.parameters() if p.requires_grad)
        self.num_params += self.unetmodel.num_params

        self.building_extractor, _, _ = load_checkpoint(epoch=30, device=device, path=dda_checkpoint)
        self._pack_cache: Dict[str, tuple] = {}

    # ------------------------------------------------------------------------------------------
    # packed-weight caches (re-folded whenever the underlying tensors change)
    # ------------------------------------------------------------------------------------------
    def _dda_pack(self, copy: str) -> torch.Tensor:
        mod = getattr(self, copy)
        tensors = list(mod.state_dict(keep_vars=True).items())
        sig = tuple((t.data_ptr(), t._version) for _, t in tensors)
        hit = self._pack_cache.get(copy)
        if hit is not None and hit[0] == sig:
            return hit[1]
        sd = {f"{copy}.{k}": t for k, t in tensors}
        dev = tensors[0][1].device
        pack = weights.pack_dda(sd, copy).to(dev)
        self._pack_cache[copy] = (sig, pack)
        return pack

    def _head_params(self):
        return tuple(getattr(self.head[i], t) for i in (0, 2, 4, 6) for t in ("weight", "bias"))

    def _head_pack(self, tc: bool = False) -> torch.Tensor:
        ps = self._head_params()
        sig = tuple((t.data_ptr(), t._version) for t in ps)
        key = "head_tc" if tc else "head"
        hit = self._pack_cache.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        with torch.no_grad():
            sd = {f"head.{i}.{t}": getattr(self.head[i], t) for i in (0, 2, 4, 6) for t in ("weight", "bias")}
            pack = weights.pack_head_tc(sd) if tc else weights.pack_head(sd)
        self._pack_cache[key] = (sig, pack)
        return pack

    # ------------------------------------------------------------------------------------------
    # reference helpers kept by name
    # ------------------------------------------------------------------------------------------
    def feature_padding(self, H: int, W: int, force: bool):
        """(top, bottom, left, right) reflect pads of add_padding (popcorn.py:231-258)."""
        if force:
            return self.p, self.p, self.p, self.p
        top = bot = left = right = 0
        if H % 32 != 0:
            t = 64 - H % 64
            top, bot = t // 2, t - t // 2
        if W % 32 != 0:
            t = 64 - W % 64
            left, right = t // 2, t - t // 2
        return top, bot, left, right

    def create_building_score(self, inputs: dict) -> torch.Tensor:
        """popcorn.py:279-322 — reflect-14, building_extractor, fusion logits, sigmoid, crop: one fused pass."""
        x = inputs["input"]
        if x.dim() != 4:
            raise ValueError("Input tensor must have shape (batch_size, channels, height, width)")
        self.unetmodel.freeze_bn_layers()
        with torch.no_grad():
            return ops.dda_forward(self._dda_pack("building_extractor"), x, self.p2d, ops.PC_DDA_BUILTUP)

    def get_sparsity_mask(self, inputs: dict, sparse_unet=False):
        """popcorn.py:325-377 (live branch).  The 60x60 grid is drawn on the host with the reference's exact CPU-RNG
        calls (:367-368) so seeded runs stay stream-compatible; mask + row-major compaction run on the GPU.
        Returns (mask bool [B,H,W], None) like the reference; the compacted index list is cached for forward()."""
        if sparse_unet:
            raise NotImplementedError("sparse_unet branch is dead code in the reference (never passed by its callers)")
        admin = inputs["admin_mask"]
        B, H, W = admin.shape
        sub = 60
        xind = torch.ones(H).multinomial(num_samples=min(sub, H), replacement=False).sort()[0]
        yind = torch.ones(W).multinomial(num_samples=min(sub, W), replacement=False).sort()[0]
        rows = torch.zeros(H, dtype=torch.uint8)
        cols = torch.zeros(W, dtype=torch.uint8)
        rows[xind] = 1
        cols[yind] = 1
        dev = admin.device
        cidx = inputs["census_idx"].to(device=dev, dtype=torch.int32).contiguous()
        bu = inputs["building_counts"][:, 0] if self.occupancymodel else None
        mask, idx, n = ops.sparse_mask_compact(bu, admin, cidx, rows.to(dev), cols.to(dev), use_builtup=self.occupancymodel)
        self._last_compaction = (idx, n)
        return mask.bool(), None

    # ------------------------------------------------------------------------------------------
    def forward(self, inputs, train=False, padding=True, return_features=True,
                encoder_no_grad=False, unet_no_grad=False, sparse=False):
        """See model/popcorn.py:100-193.  ``train`` and ``return_features`` are accepted and unused, as there.

        Region ids: ``admin_mask`` holds integral ids (float32 in the reference's samples, data/PopulationDataset.py:445; -1 = collate
        padding, 0 = background) that float32 represents exactly, i.e. |id| <= 2**24, and ``census_idx`` integers in the same range.
        Every path compares them as the reference does (``admin_mask == census_idx``): the kernels take int32 / float32 copies of these
        exactly-representable values, so the dense, sparse and autograd paths select the same pixels.  POPCORN_CHECK_IDS=1 verifies
        the precondition (one pass + host sync) and raises ValueError on fractional or out-of-range ids."""
        X = inputs["input"]
        if X.dim() != 4:
            raise ValueError("Input tensor must have shape (batch_size, channels, height, width)")
        if not X.is_cuda:
            raise RuntimeError("popcorn_b200.POPCORN runs on CUDA (sm_100a) tensors only; there is no CPU fallback")

        if _CHECK_IDS and "admin_mask" in inputs.keys():
            am = inputs["admin_mask"]
            if am.is_floating_point() and not bool(((am == am.round()) & (am.abs() <= 2 ** 24)).all()):
                raise ValueError("admin_mask must hold integral region ids with |id| <= 2**24 (float32-exact)")
        # eval fast path: builtup pass + feature pass + tcgen05 head + census partials behind ONE library call (pc_infer_tile_fused)
        B, _, H, W = X.shape
        if (USE_TENSOR_CORE_HEAD and FUSED_EVAL and not sparse and not padding and not torch.is_grad_enabled() and self.occupancymodel
                and ("building_counts" not in inputs.keys() or self.sentinelbuildings) and H >= 29 and W >= 29):
            self.unetmodel.freeze_bn_layers()
            sums = torch.zeros(B, dtype=torch.float64, device=X.device)
            if "admin_mask" in inputs.keys():
                ids = inputs["admin_mask"].to(torch.int32).contiguous()
                cidx = inputs["census_idx"].to(device=X.device, dtype=torch.int32).contiguous()
            else:
                ids, cidx = None, torch.zeros(B, dtype=torch.int32, device=X.device)       # bin = batch index, all pixels
            dens, scale, builtup = ops.infer_tile_fused(self._dda_pack("building_extractor"), self._dda_pack("unetmodel"),
                                                        self._head_pack(tc=True), X, ids, cidx, sums, want_scale=True)
            inputs["building_counts"] = builtup
            return {"popcount": sums.float(), "popdensemap": dens, "scale": scale, "builtup_score": builtup, "occupancy": scale}

        # builtup score (popcorn.py:112-115)
        if "building_counts" not in inputs.keys() or self.sentinelbuildings:
            inputs["building_counts"] = self.create_building_score(inputs)
        builtup = inputs["building_counts"]
        if builtup.dtype != torch.float32 or not builtup.is_contiguous():
            builtup = builtup.float().contiguous()

        aux = {}
        if sparse:
            sparsity_mask, _ = self.get_sparsity_mask(inputs)
            idx, n_dev = self._last_compaction

        # feature pass (popcorn.py:126-158); BN is frozen / eval on every call (:128)
        self.unetmodel.freeze_bn_layers()
        unet_trainable = any(p.requires_grad for p in self.unetmodel.parameters())
        B, _, H, W = X.shape
        pads = self.feature_padding(H, W, force=bool(padding))
        if torch.is_grad_enabled() and not unet_no_grad and unet_trainable:
            # fine-tuning path (run_train.py:191-202 for batches < 9 M px): layer-by-layer forward that keeps the
            # activations, hand-written backward (model/unet_train.py, csrc/unet_bwd.cu)
            feats = unet_train.unet_features(self.unetmodel, X, pads, encoder_no_grad, self.S1, self.S2)
        else:
            with torch.no_grad():
                feats = ops.dda_forward(self._dda_pack("unetmodel"), X, pads, ops.PC_DDA_FEATURES)

        bu = builtup if self.occupancymodel else None
        need_grad = torch.is_grad_enabled() and (feats.requires_grad or any(p.requires_grad for p in self.head.parameters()))
        has_admin = "admin_mask" in inputs.keys()

        if sparse or need_grad:
            if not sparse:   # dense head with autograd: every pixel is "selected"
                idx = torch.arange(B * H * W, dtype=torch.int32, device=X.device)
                n_dev = torch.full((1,), B * H * W, dtype=torch.int32, device=X.device)
                n = B * H * W
            params = self._head_params()
            with torch.no_grad():          # weight packs first: their launches queue behind the DDA passes while the host waits for n
                hsd = {f"head.{i}.{t}": p for (i, t), p in zip(((0, "weight"), (0, "bias"), (2, "weight"), (2, "bias"),
                                                               (4, "weight"), (4, "bias"), (6, "weight"), (6, "bias")), params)}
                hpack = weights.pack_head(hsd)                       # fp32 pack: the backward recomputes the forward from it
                tcpack = weights.pack_head_tc(hsd) if USE_TENSOR_CORE_HEAD else None
            if sparse:
                n = int(n_dev.item())      # the reference's boolean indexing synchronises here as well
            if need_grad:
                pop_sel, dens, scale_sel = _SparseHeadFn.apply(feats, bu, idx, n_dev, n, self.head_input_dim, hpack, tcpack, *params)
            else:
                with torch.no_grad():
                    pop_sel, dens, scale_sel = _SparseHeadFn.apply(feats, bu, idx, n_dev, n, self.head_input_dim, hpack, tcpack, *params)
            popdensemap = dens
            if self.occupancymodel:
                aux["scale"] = scale_sel if sparse else scale_sel.view(B, H, W)
            else:
                aux["scale"] = None
            if sparse:
                # the mask lies inside the region, so the masked sum equals the sum over selected pixels (:186-187)
                popcount = pop_sel
            elif has_admin:
                this_mask = inputs["admin_mask"] == inputs["census_idx"].view(-1, 1, 1)
                popcount = (popdensemap * this_mask).sum((1, 2))
            else:
                popcount = pop_sel
        else:
            hpack = self._head_pack(tc=USE_TENSOR_CORE_HEAD)
            sums = torch.zeros(B, dtype=torch.float64, device=X.device)
            ids = cidx = None
            if has_admin:
                ids = inputs["admin_mask"].to(torch.int32).contiguous()
                cidx = inputs["census_idx"].to(device=X.device, dtype=torch.int32).contiguous()
            else:
                cidx = torch.zeros(B, dtype=torch.int32, device=X.device)   # bin = batch index, all pixels
            dens, scale = ops.head_dense_forward(hpack, feats, bu, ids, cidx, sums, want_scale=self.occupancymodel,
                                                 tc=USE_TENSOR_CORE_HEAD)
            popdensemap = dens
            aux["scale"] = scale if self.occupancymodel else None
            popcount = sums.float()

        out = {"popcount": popcount, "popdensemap": popdensemap, **aux}
        # extra aliases named by BASELINE.json's north_star (not present in the reference dict)
        out["builtup_score"] = builtup
        out["occupancy"] = aux.get("scale")
        return out

    # kept for API compatibility with callers that poke at the reference helpers
    def add_padding(self, data: torch.Tensor, force=True):
        """Materialising variant of popcorn.py:231-258 (the kernels pad virtually; this is only for callers)."""
        H, W = data.shape[2:]
        top, bot, left, right = self.feature_padding(H, W, force)
        if top or bot:
            data = nn.functional.pad(data, (0, 0, top, bot), mode="reflect")
        if left or right:
            data = nn.functional.pad(data, (left, right, 0, 0), mode="reflect")
        none = lambda a, b: (None, None) if (a == 0 and b == 0) else (a, b)
        px1, px2 = none(top, bot)
        py1, py2 = none(left, right)
        return data, (px1, px2, py1, py2)

    def revert_padding(self, data: torch.Tensor, padding: tuple):
        """popcorn.py:261-276."""
        px1, px2, py1, py2 = padding
        if px1 is not None or px2 is not None:
            data = data[:, :, px1:-px2, :]
        if py1 is not None or py2 is not None:
            data = data[:, :, :, py1:-py2]
        return data
