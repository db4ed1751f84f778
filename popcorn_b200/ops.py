"""Thin torch-tensor front ends over the C-ABI (pointers, strides, stream, workspace)."""
from __future__ import annotations

import functools
from typing import Optional

import torch

from . import _lib

PC_DDA_FEATURES, PC_DDA_BUILTUP = 0, 1


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    """Every tensor of a call must live on ONE CUDA device; returns that device.  (There is no CPU path; kernels launched with
    pointers of another device's context would fault or, worse, silently read foreign memory.)"""
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("popcorn_b200: tensors must live on a CUDA device (no CPU path exists)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"popcorn_b200: tensors of one call live on different devices ({dev} and {t.device})")
    return dev


class _on:
    """Make the tensors' device current for the duration of a call: the library launches on torch's CURRENT stream of the CURRENT
    device, so a model built on cuda:1 must not launch while cuda:0 is current."""

    def __init__(self, dev):
        self.guard = torch.cuda.device(dev) if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device() else None

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()

    def __exit__(self, *a):
        if self.guard is not None:
            self.guard.__exit__(*a)


def _device_guard(fn):
    """Run `fn` with the device of its first CUDA tensor argument current (see _on)."""
    def first_cuda(objs):
        for o in objs:
            if isinstance(o, torch.Tensor):
                if o.is_cuda:
                    return o.device
            elif isinstance(o, (tuple, list)):
                d = first_cuda(o)
                if d is not None:
                    return d
        return None

    @functools.wraps(fn)
    def wrapper(*a, **k):
        with _on(first_cuda(list(a) + list(k.values()))):
            return fn(*a, **k)
    return wrapper


class Workspace:
    """Grow-only scratch owned by torch's caching allocator, one buffer per (device, stream): work on different streams (or
    devices, or autograd's backward thread on another stream) never shares scratch.  Within one stream launches are ordered, so the
    DDA activations, the compaction counters and the backward partials can re-use the same bytes one after the other.  A buffer
    that is outgrown goes back to the allocator, which is stream-aware: the block is not handed to another stream while work
    queued on this one may still touch it (record_stream below marks it as used by the launching stream)."""

    def __init__(self):
        self.bufs = {}

    def get(self, nbytes: int, device) -> torch.Tensor:
        device = torch.device(device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        stream = torch.cuda.current_stream(idx)
        key = (idx, stream.cuda_stream)
        buf = self.bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes), dtype=torch.uint8, device=torch.device("cuda", idx))
            buf.record_stream(stream)
            self.bufs[key] = buf
        return buf


_ws = Workspace()


@_device_guard
def dda_forward(wpack: torch.Tensor, x: torch.Tensor, pads=(0, 0, 0, 0), mode: int = PC_DDA_FEATURES,
                out: Optional[torch.Tensor] = None, workspace: Optional[Workspace] = None) -> torch.Tensor:
    """x [B,C,H,W] fp32 (any strides with unit W stride) -> features [B,F,H,W] or builtup score [B,1,H,W]."""
    _need_cuda(wpack, x)
    L = _lib.lib()
    if x.dim() != 4:
        raise ValueError("Input tensor must have shape (batch_size, channels, height, width)")
    if x.dtype != torch.float32 or x.stride(3) != 1:
        x = x.float().contiguous()
    B, Cc, H, W = x.shape
    top, bot, left, right = pads
    nf = 16 if Cc == 6 else 8
    oc = nf if mode == PC_DDA_FEATURES else 1
    if out is None:
        out = torch.empty(B, oc, H, W, dtype=torch.float32, device=x.device)
    ws = (workspace or _ws)
    need = L.pc_dda_workspace_bytes(B, Cc, H + top + bot, W + left + right)
    buf = ws.get(need, x.device)
    _lib.check(L.pc_dda_forward(wpack.data_ptr(), wpack.numel(), x.data_ptr(), B, Cc, H, W, x.stride(0), x.stride(1), x.stride(2),
                                top, bot, left, right, mode, out.data_ptr(), out.stride(0), out.stride(1),
                                out.stride(2), buf.data_ptr(), buf.numel(), _stream()), "pc_dda_forward")
    return out


@_device_guard
def infer_tile_fused(bext_pack: torch.Tensor, unet_pack: torch.Tensor, head_tcpack: torch.Tensor, x: torch.Tensor,
                     ids: Optional[torch.Tensor] = None, census_idx: Optional[torch.Tensor] = None,
                     sums: Optional[torch.Tensor] = None, want_scale: bool = True):
    """Dense eval forward of a tile batch in one library call (pc_infer_tile_fused): -> (dens [B,H,W], scale [B,H,W]|None,
    builtup [B,1,H,W]); `sums` (float64) is updated in place as in head_dense_forward."""
    _need_cuda(bext_pack, unet_pack, head_tcpack, x, ids, census_idx, sums)
    L = _lib.lib()
    if x.dim() != 4:
        raise ValueError("Input tensor must have shape (batch_size, channels, height, width)")
    if x.dtype != torch.float32 or x.stride(3) != 1:
        x = x.float().contiguous()
    B, Cc, H, W = x.shape
    assert bext_pack.numel() == unet_pack.numel()
    dens = torch.empty(B, H, W, dtype=torch.float32, device=x.device)
    scale = torch.empty_like(dens) if want_scale else None
    builtup = torch.empty(B, 1, H, W, dtype=torch.float32, device=x.device)
    if ids is not None:
        assert ids.dtype == torch.int32 and ids.is_contiguous() and tuple(ids.shape) == (B, H, W)
    if sums is not None:
        assert sums.dtype == torch.float64 and sums.is_contiguous()
    need = L.pc_infer_tile_workspace_bytes(B, Cc, H, W)
    buf = _ws.get(need, x.device)
    _lib.check(L.pc_infer_tile_fused(bext_pack.data_ptr(), unet_pack.data_ptr(), bext_pack.numel(), head_tcpack.data_ptr(),
                                     x.data_ptr(), B, Cc, H, W, x.stride(0), x.stride(1), x.stride(2), dens.data_ptr(), _ptr(scale),
                                     builtup.data_ptr(), _ptr(ids), _ptr(census_idx), _ptr(sums), 0 if sums is None else sums.numel(),
                                     buf.data_ptr(), buf.numel(), _stream()), "pc_infer_tile_fused")
    return dens, scale, builtup


@_device_guard
def head_dense_forward(hpack, feats, builtup, ids=None, census_idx=None, sums=None, want_scale=True, tc=False):
    """feats [B,Cin,H,W], builtup [B,1,H,W]|None -> (dens [B,H,W], scale [B,H,W]|None); sums (float64) updated in place.
    tc=True: `hpack` is the tcgen05 weight image (weights.pack_head_tc) and the tensor-core kernel runs."""
    _need_cuda(hpack, feats, builtup, ids, census_idx, sums)
    L = _lib.lib()
    fn = L.pc_head_dense_forward_tc if tc else L.pc_head_dense_forward
    B, Cin, H, W = feats.shape
    assert feats.stride(3) == 1 and feats.dtype == torch.float32
    dens = torch.empty(B, H, W, dtype=torch.float32, device=feats.device)
    scale = torch.empty_like(dens) if want_scale else None
    if builtup is not None:
        assert builtup.stride(-1) == 1 and builtup.dtype == torch.float32
    if ids is not None:
        assert ids.dtype == torch.int32 and ids.stride(-1) == 1 and ids.dim() == 3
    if census_idx is not None:
        assert census_idx.dtype == torch.int32
    R = 0 if sums is None else sums.numel()
    if sums is not None:
        assert sums.dtype == torch.float64 and sums.is_contiguous()
    _lib.check(fn(
        hpack.data_ptr(), Cin, feats.data_ptr(), feats.stride(0), feats.stride(1), feats.stride(2),
        _ptr(builtup), 0 if builtup is None else builtup.stride(0), 0 if builtup is None else builtup.stride(-2),
        B, H, W, dens.data_ptr(), _ptr(scale), dens.stride(0), dens.stride(1),
        _ptr(ids), 0 if ids is None else ids.stride(0), 0 if ids is None else ids.stride(1),
        _ptr(census_idx), _ptr(sums), R, _stream()), "pc_head_dense_forward")
    return dens, scale


@_device_guard
def sparse_mask_compact(builtup, admin, census_idx, grid_rows, grid_cols, use_builtup=True):
    """-> (mask uint8 [B,H,W], idx int32 [B*H*W] (first n valid), n int32[1] device)."""
    _need_cuda(builtup, admin, census_idx, grid_rows, grid_cols)
    L = _lib.lib()
    B, H, W = admin.shape
    admin = admin.float().contiguous()
    bu = None if builtup is None else builtup.float().contiguous()
    dev = admin.device
    mask = torch.empty(B, H, W, dtype=torch.uint8, device=dev)
    idx = torch.empty(B * H * W, dtype=torch.int32, device=dev)
    n = torch.zeros(1, dtype=torch.int32, device=dev)
    need = L.pc_compact_workspace_bytes(B * H * W)
    buf = _ws.get(need, dev)
    _lib.check(L.pc_sparse_mask_compact(_ptr(bu), admin.data_ptr(), census_idx.data_ptr(), grid_rows.data_ptr(),
                                        grid_cols.data_ptr(), 1 if use_builtup else 0, B, H, W, mask.data_ptr(),
                                        idx.data_ptr(), n.data_ptr(), buf.data_ptr(), buf.numel(), _stream()),
               "pc_sparse_mask_compact")
    return mask, idx, n


@_device_guard
def head_sparse_forward(hpack, feats, builtup, idx, n_dev, n_max, tc=False):
    """-> (dens [B,H,W] scattered, scale_sel [n_max] (first n valid), popcount float64 [B])."""
    _need_cuda(hpack, feats, builtup, idx, n_dev)
    L = _lib.lib()
    fn = L.pc_head_sparse_forward_tc if tc else L.pc_head_sparse_forward
    B, Cin, H, W = feats.shape
    assert feats.is_contiguous() and feats.dtype == torch.float32
    bu = None if builtup is None else builtup.float().contiguous()
    dens = torch.zeros(B, H, W, dtype=torch.float32, device=feats.device)
    scale_sel = torch.empty(max(int(n_max), 1), dtype=torch.float32, device=feats.device)
    pop = torch.zeros(B, dtype=torch.float64, device=feats.device)
    _lib.check(fn(hpack.data_ptr(), Cin, feats.data_ptr(), feats.stride(0), feats.stride(1),
                                        _ptr(bu), idx.data_ptr(), n_dev.data_ptr(), int(n_max), H * W, dens.data_ptr(),
                                        scale_sel.data_ptr(), pop.data_ptr(), _stream()), "pc_head_sparse_forward")
    return dens, scale_sel, pop


@_device_guard
def head_sparse_backward(hpack, feats, builtup, idx, n_dev, n_max, g_pop, g_coef, g_sel=None, want_g_feats=False):
    """-> gradient buffer in hpack layout (fp32) [, dL/dfeats [B,Cin,H,W] when want_g_feats]."""
    _need_cuda(hpack, feats, builtup, idx, n_dev, g_pop, g_sel)
    L = _lib.lib()
    B, Cin, H, W = feats.shape
    bu = None if builtup is None else builtup.float().contiguous()
    grad = torch.empty_like(hpack)
    need = L.pc_head_bwd_workspace_bytes(Cin)
    buf = _ws.get(need, feats.device)
    g_pop = g_pop.float().contiguous()
    if g_sel is not None:
        g_sel = g_sel.float().contiguous()
    g_feats = torch.zeros_like(feats) if want_g_feats else None
    _lib.check(L.pc_head_sparse_backward(hpack.data_ptr(), Cin, feats.data_ptr(), feats.stride(0), feats.stride(1),
                                         _ptr(bu), idx.data_ptr(), n_dev.data_ptr(), int(n_max), H * W,
                                         g_pop.data_ptr(), float(g_coef), _ptr(g_sel), grad.data_ptr(),
                                         buf.data_ptr(), buf.numel(), _ptr(g_feats), _stream()), "pc_head_sparse_backward")
    return (grad, g_feats) if want_g_feats else grad


@_device_guard
def region_sum(dens: torch.Tensor, ids: torch.Tensor, R: int, sums: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sums[id] += dens (float64 [R]); ids int32, same shape as dens, both contiguous."""
    _need_cuda(dens, ids)
    assert dens.is_contiguous() and ids.is_contiguous() and ids.dtype == torch.int32 and dens.dtype == torch.float32
    assert dens.numel() == ids.numel()
    if sums is None:
        sums = torch.zeros(R, dtype=torch.float64, device=dens.device)
    if dens.numel() == 0:
        return sums
    _lib.check(_lib.lib().pc_region_sum(dens.data_ptr(), ids.data_ptr(), dens.numel(), R, sums.data_ptr(), _stream()),
               "pc_region_sum")
    return sums


@_device_guard
def region_sum_backward(g_sums: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    _need_cuda(g_sums, ids)
    g = g_sums.float().contiguous()
    out = torch.empty(ids.shape, dtype=torch.float32, device=ids.device)
    _lib.check(_lib.lib().pc_region_sum_backward(g.data_ptr(), ids.data_ptr(), ids.numel(), g.numel(), out.data_ptr(),
                                                 _stream()), "pc_region_sum_backward")
    return out


@_device_guard
def region_scale_(dens: torch.Tensor, ids: torch.Tensor, factor: torch.Tensor) -> torch.Tensor:
    _need_cuda(dens, ids, factor)
    assert dens.is_contiguous() and ids.is_contiguous() and factor.dtype == torch.float32
    _lib.check(_lib.lib().pc_region_scale(dens.data_ptr(), ids.data_ptr(), dens.numel(), factor.numel(),
                                          factor.data_ptr(), _stream()), "pc_region_scale")
    return dens


@_device_guard
def accumulate_tile(dens, scale, rows, cols, maps, y0, x0):
    """maps = (map, map_sq, smap, smap_sq, count) full-raster tensors (any may be None except map)."""
    m, msq, sm, ssq, cnt = maps
    _lib.check(_lib.lib().pc_accumulate_tile(dens.data_ptr(), _ptr(scale), dens.stride(-2), rows[0], rows[1], cols[0],
                                             cols[1], m.data_ptr(), _ptr(msq), _ptr(sm), _ptr(ssq), _ptr(cnt),
                                             m.stride(0), y0, x0, _stream()), "pc_accumulate_tile")


@_device_guard
def finalize_map(maps, rows=None):
    """Mean / std where a pixel was visited more than once (run_eval.py:140-154); ``rows`` = (r0, r1) restricts it to a
    row range of the maps (used to finalise and ship finished strips while later strips still compute)."""
    m, msq, sm, ssq, cnt = maps
    if m.numel() == 0:          # a rank that owns no rows (more ranks than row strips): nothing to finalise
        return
    if rows is not None:
        r0, r1 = rows
        if r1 <= r0:
            return
        m, cnt = m[r0:r1], cnt[r0:r1]
        msq = None if msq is None else msq[r0:r1]
        sm = None if sm is None else sm[r0:r1]
        ssq = None if ssq is None else ssq[r0:r1]
    _lib.check(_lib.lib().pc_finalize_map(m.data_ptr(), _ptr(msq), _ptr(sm), _ptr(ssq), cnt.data_ptr(), m.numel(),
                                          _stream()), "pc_finalize_map")


# per-band normalisation statistics of the reference (data/config/dataset_stats.json: "sen2springNIR", "sen2spring", "sen1"),
# in the reference's channel order S2 = (R, G, B[, NIR]), S1 = (VV, VH)
DATASET_STATS = {
    "sen2springNIR": {"mean": (1460.4567, 1468.2986, 1383.4556, 2226.6821), "std": (1130.7949, 1129.0261, 1053.3217, 1724.3213)},
    "sen2spring": {"mean": (1460.4567, 1468.2986, 1383.4556), "std": (1130.7949, 1129.0261, 1053.3217)},
    "sen1": {"mean": (-11.426, -17.753), "std": (5.5983, 5.0076)},
}
S2_FILE_TO_RGBN = 0x03000102      # GeoTIFF band order B02,B03,B04,B08 -> R,G,B,NIR (S2_RGBNIR_channels = (3,2,1,4), PopulationDataset.py:566)
S2_IDENTITY = 0x03020100


@_device_guard
def ingest_normalize(s2: Optional[torch.Tensor], s1: Optional[torch.Tensor], out: Optional[torch.Tensor] = None,
                     s2_plane_map: int = S2_IDENTITY, stats: Optional[dict] = None, stream: Optional[int] = None) -> torch.Tensor:
    """Raw bands on the DEVICE -> normalised fp32 window [n_s2+n_s1, h, w] = cat[(S2-mean)/std, (S1-mean)/std]
    (utils/utils.py:105-127, 162-171).  s2: [3|4,h,w] uint16 or float32 view (unit column stride); s1: [2,h,w] float32."""
    import ctypes as C
    _need_cuda(s2, s1, out)
    stats = stats or DATASET_STATS
    n2 = 0 if s2 is None else s2.shape[0]
    n1 = 0 if s1 is None else s1.shape[0]
    ref = s2 if s2 is not None else s1
    h, w = ref.shape[1], ref.shape[2]
    mean, std = [], []
    if n2:
        assert s2.dtype in (torch.uint16, torch.int16, torch.float32) and s2.stride(2) == 1
        key = "sen2springNIR" if n2 == 4 else "sen2spring"
        mean += list(stats[key]["mean"]); std += list(stats[key]["std"])
    if n1:
        assert s1.dtype == torch.float32 and s1.stride(2) == 1 and s1.shape[1:] == ref.shape[1:]
        mean += list(stats["sen1"]["mean"]); std += list(stats["sen1"]["std"])
    if out is None:
        out = torch.empty(n2 + n1, h, w, dtype=torch.float32, device=ref.device)
    assert out.shape == (n2 + n1, h, w) and out.stride(2) == 1 and out.dtype == torch.float32
    m = (C.c_float * 6)(*mean)
    s = (C.c_float * 6)(*std)
    _lib.check(_lib.lib().pc_ingest_normalize(
        _ptr(s2), 1 if (n2 and s2.dtype != torch.float32) else 0, n2, 0 if not n2 else s2.stride(0), 0 if not n2 else s2.stride(1),
        s2_plane_map, _ptr(s1), n1, 0 if not n1 else s1.stride(0), 0 if not n1 else s1.stride(1), h, w, m, s,
        out.data_ptr(), out.stride(0), out.stride(1), _stream() if stream is None else stream), "pc_ingest_normalize")
    return out


def launch_count(reset: bool = False) -> int:
    """Kernels launched by libpopcorn_b200 so far in this process."""
    return int(_lib.lib().pc_launch_count(1 if reset else 0))


@_device_guard
def copy_window_h2d(dst: torch.Tensor, src: torch.Tensor, stream: Optional[int] = None) -> None:
    """dst [C,h,w] device (contiguous rows) <- src [C,h,w] view of a pinned host raster (unit column stride)."""
    assert dst.is_cuda and not src.is_cuda and src.stride(2) == 1 and dst.stride(2) == 1 and dst.shape == src.shape
    L = _lib.lib()
    st = _stream() if stream is None else stream
    C, h, w = src.shape
    es = src.element_size()
    for c in range(C):
        _lib.check(L.pc_memcpy2d_async(dst[c].data_ptr(), dst.stride(1) * es, src[c].data_ptr(), src.stride(1) * es,
                                       w * es, h, 1, st), "pc_memcpy2d_async")


@_device_guard
def copy_d2h(dst: torch.Tensor, src: torch.Tensor, stream: Optional[int] = None) -> None:
    """dst pinned host [h,w] <- src device [h,w] (both unit column stride)."""
    assert src.is_cuda and not dst.is_cuda and dst.shape == src.shape and src.dim() == 2
    es = src.element_size()
    _lib.check(_lib.lib().pc_memcpy2d_async(dst.data_ptr(), dst.stride(0) * es, src.data_ptr(), src.stride(0) * es,
                                            src.shape[1] * es, src.shape[0], 2, _stream() if stream is None else stream),
               "pc_memcpy2d_async")


def profile_enable(on: bool) -> None:
    """Bracket every library launch with CUDA events (on its own stream) until disabled."""
    _lib.lib().pc_profile_enable(1 if on else 0)


def profile_results() -> dict:
    """{kernel name: (total ms, launches, pixels processed)} for the launches since the last profile_enable(True)."""
    import ctypes as C
    L = _lib.lib()
    out = {}
    for i in range(L.pc_profile_num()):
        ms, n, u = C.c_double(0), C.c_longlong(0), C.c_double(0)
        _lib.check(L.pc_profile_get(i, C.byref(ms), C.byref(n), C.byref(u)), "pc_profile_get")
        if n.value:
            out[L.pc_profile_name(i).decode()] = (ms.value, n.value, u.value)
    return out
