"""A CPU stand-in for the device side of popcorn_b200.country / timeseries — TEST INFRASTRUCTURE for the host logic only.

The engine's window arithmetic (row offsets of sharded rasters, upload rings, chunked slabs, strip shipping, empty ranks)
is plain Python around `ops.*` calls and `torch.cuda` streams/events.  `install(monkeypatch)` swaps, inside those two
modules only, `torch` for a proxy whose "cuda" device is the CPU and whose streams/events are no-ops, and `ops` for
torch-CPU functions with the same signatures and a *3x3-local* fake network (so halos matter by one pixel, like the real
23-pixel receptive field matters against the 128-pixel overlap).  Nothing here is reachable from the product: the real
path has no CPU implementation, and the GPU tests cover the same engine on the kernels.
"""
from __future__ import annotations

import contextlib
import types

import torch
import torch.nn.functional as F

from popcorn_b200 import ops as real_ops

S2_FILE_TO_RGBN = real_ops.S2_FILE_TO_RGBN
STATS = real_ops.DATASET_STATS


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, *a, **k):
        pass


class _Stream:
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass

    def wait_event(self, ev):
        pass

    def synchronize(self):
        pass


class _FakeCuda:
    Event, Stream = _Event, _Stream

    @staticmethod
    def current_device():
        return 0

    @staticmethod
    def current_stream(dev=None):
        return _Stream()

    @staticmethod
    def stream(s):
        return contextlib.nullcontext()

    @staticmethod
    def synchronize():
        pass


class _TorchProxy:
    """`torch` with a fake `.cuda` namespace and `torch.device("cuda", i)` -> cpu."""
    cuda = _FakeCuda

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def device(*a, **k):
        return torch.device("cpu")

    @staticmethod
    def _fix(kw):
        if "device" in kw:
            kw["device"] = "cpu"
        return kw

    def zeros(self, *a, **k):
        return torch.zeros(*a, **self._fix(k))

    def empty(self, *a, **k):
        return torch.empty(*a, **self._fix(k))


# ---- the fake network: pointwise except for one 3x3 box filter --------------------------------------------------
def _box(x):
    return F.avg_pool2d(x, 3, 1, 1, count_include_pad=True)


def fake_features(x):            # [B,6,H,W] -> [B,16,H,W]
    return torch.cat([_box(x), x, x[:, :4]], 1)


def fake_builtup(x):             # -> [B,1,H,W] in (0,1)
    return torch.sigmoid(_box(x).sum(1, keepdim=True))


def fake_density(x):
    scale = torch.relu(fake_features(x).sum(1))
    return scale * fake_builtup(x)[:, 0], scale


def normalise(s2, s1, plane_map=S2_FILE_TO_RGBN, stats=None):
    stats = stats or STATS
    order = [(plane_map >> (8 * c)) & 0xFF for c in range(4)]
    s2 = s2[order].to(torch.float32)
    m2, d2 = torch.tensor(stats["sen2springNIR"]["mean"]).view(4, 1, 1), torch.tensor(stats["sen2springNIR"]["std"]).view(4, 1, 1)
    m1, d1 = torch.tensor(stats["sen1"]["mean"]).view(2, 1, 1), torch.tensor(stats["sen1"]["std"]).view(2, 1, 1)
    return torch.cat([(s2 - m2) / d2, (s1 - m1) / d1], 0)


class FakeOps(types.SimpleNamespace):
    pass


def make_ops(log):
    o = FakeOps(PC_DDA_FEATURES=0, PC_DDA_BUILTUP=1, S2_FILE_TO_RGBN=S2_FILE_TO_RGBN, S2_IDENTITY=real_ops.S2_IDENTITY,
                DATASET_STATS=STATS)

    def dda_forward(wpack, x, pads=(0, 0, 0, 0), mode=0, **kw):
        log.append(("dda", mode, tuple(x.shape)))
        return fake_builtup(x) if mode == 1 else fake_features(x)

    def head_dense_forward(hpack, feats, builtup, ids=None, census_idx=None, sums=None, want_scale=True, tc=False):
        scale = torch.relu(feats.sum(1))
        dens = scale * builtup[:, 0] if builtup is not None else scale
        return dens, (scale if want_scale else None)

    def accumulate_tile(dens, scale, rows, cols, maps, y0, x0):
        m, msq, sm, ssq, cnt = maps
        r0, r1 = rows
        c0, c1 = cols
        tgt = (slice(y0 + r0, y0 + r1), slice(x0 + c0, x0 + c1))
        assert y0 + r0 >= 0 and y0 + r1 <= m.shape[0], "window centre outside the rows this rank owns"
        d = dens[r0:r1, c0:c1]
        m[tgt] += d
        if msq is not None:
            msq[tgt] += d * d
        if scale is not None and sm is not None:
            sm[tgt] += scale[r0:r1, c0:c1]
        if scale is not None and ssq is not None:
            ssq[tgt] += scale[r0:r1, c0:c1] ** 2
        cnt[tgt] += 1

    def finalize_map(maps, rows=None):
        m, msq, sm, ssq, cnt = maps
        if m.numel() == 0:
            return
        r0, r1 = rows if rows is not None else (0, m.shape[0])
        if r1 <= r0:
            return
        log.append(("finalize", r0, r1))
        sl = slice(r0, r1)
        n = cnt[sl].float()
        multi = cnt[sl] > 1
        for a, sq in ((m, msq), (sm, ssq)):
            if a is None:
                continue
            mean = torch.where(multi, a[sl] / n, a[sl])
            if sq is not None:
                sq[sl] = torch.where(multi, ((sq[sl] - mean * mean * n) / (n - 1)).clamp_min(0).sqrt(), sq[sl])
            a[sl] = mean

    def region_sum(dens, ids, R, sums=None):
        if sums is None:
            sums = torch.zeros(R, dtype=torch.float64)
        if dens.numel():
            sums.index_add_(0, ids.reshape(-1).long(), dens.reshape(-1).double())
        return sums

    def copy_window_h2d(dst, src, stream=None):
        assert dst.shape == src.shape
        log.append(("h2d", src.numel() * src.element_size()))
        dst.copy_(src)

    def copy_d2h(dst, src, stream=None):
        assert dst.shape == src.shape
        log.append(("d2h", tuple(src.shape)))
        dst.copy_(src)

    def ingest_normalize(s2, s1, out=None, s2_plane_map=real_ops.S2_IDENTITY, stats=None, stream=None):
        v = normalise(s2, s1, s2_plane_map, stats)
        if out is None:
            return v
        out.copy_(v)
        return out

    for f in (dda_forward, head_dense_forward, accumulate_tile, finalize_map, region_sum, copy_window_h2d, copy_d2h,
              ingest_normalize):
        setattr(o, f.__name__, f)
    return o


class FakeModel:
    occupancymodel = True
    p2d = (14, 14, 14, 14)

    def _dda_pack(self, copy):
        return torch.zeros(1)

    def _head_pack(self, tc=False):
        return torch.zeros(1)


def install(monkeypatch):
    """Patch popcorn_b200.country / timeseries for the duration of a test; returns the call log."""
    from popcorn_b200 import country as ct
    from popcorn_b200 import timeseries as ts
    log = []
    proxy = _TorchProxy()
    fops = make_ops(log)
    for mod in (ct, ts):
        monkeypatch.setattr(mod, "torch", proxy)
    monkeypatch.setattr(ct, "ops", fops)
    monkeypatch.setattr(ct.RawRaster, "is_pinned", lambda self: True)
    return log
