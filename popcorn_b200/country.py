"""Country-scale tiled inference + census aggregation, sharded over the GPUs of one box.

Restates the reference's evaluation loop (run_eval.py:83-154: 2048^2 tiles, 128-px overlap, centre-only
write-back, visit-count averaging; data/PopulationDataset.py:294-334 tile grid, :656-672 centre mask,
:696-729 census sums, :823-852 dasymetric adjustment) on top of the sm_100a kernels:

* ``plan_windows``  — the reference tile grid, optionally *merged* into row strips.  Main-grid tiles abut
  exactly (stride 1792 = 2048 - 2*128) and their origins are multiples of 4, so a merged window reproduces
  every tile's pool phase; results on the written centre pixels are bit-identical to per-tile execution
  while the halo recompute drops from 1.31x to ~1.07x.  Edge tiles (bottom row / right column / corner,
  origin h-2048 / w-2048) keep their own phase and are merged only among themselves.
* ``shard_windows`` — contiguous tile-row blocks per rank; the bottom edge row goes to the owner of the last
  main tile-row so that both contributors of any doubly-covered pixel live on one GPU.
* ``CountryEngine`` — runs the windows (device-resident raster, or pinned-host raster streamed through a
  double-buffered copy stream), accumulates sum / sum-of-squares / count maps on the device, finalises
  mean / std, segment-sums the map over the census id raster and all-reduces the R partial sums (NCCL).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import ops
from .model import popcorn as popcorn_mod

PATCH = 2048      # utils/constants.py:12
OVERLAP = 128     # utils/constants.py:13
RF_RADIUS = 24    # receptive-field radius of the DDA UNet is 23 px -> 24 keeps the 4-px phase


def can_merge(patch: int, overlap: int) -> bool:
    """Merged windows equal per-tile execution on the written centre only if the discarded frame (overlap) covers the
    receptive field of the tile border (zero / reflect padding artefacts) and the tile stride keeps the pool phase."""
    return overlap >= RF_RADIUS and (patch - 2 * overlap) % 4 == 0 and patch % 4 == 0


@dataclass(frozen=True)
class Window:
    """Input window [y0:y0+h, x0:x0+w] of the raster; its centre [ov:-ov] is written back.
    ``tile_row`` = index of the first main-grid tile-row it covers (-1 for the bottom edge row)."""
    y0: int
    x0: int
    h: int
    w: int
    tile_row: int
    ntiles: int


def grid_origins(n: int, patch: int = PATCH, overlap: int = OVERLAP) -> List[int]:
    """Main-grid origins along one axis: arange(0, n - patch, patch - 2*overlap) (PopulationDataset.py:301-305)."""
    return list(range(0, n - patch, patch - 2 * overlap))


def strip_sizes(n_rows: int, rows_per_strip: int, first_strip_rows: Optional[int] = None, last_strip_rows: Optional[int] = None) -> List[int]:
    """Tile-rows per merged strip: a short first strip (a streamed run starts computing after a small upload), `rows_per_strip` in
    the middle (long kernels, little halo recompute), a short last strip (little map left to ship when the compute ends)."""
    sizes, left = [], n_rows
    if first_strip_rows and left > 0:
        k = min(first_strip_rows, left)
        sizes.append(k)
        left -= k
    tail = min(last_strip_rows, left) if (last_strip_rows and left > 0) else 0
    left -= tail
    while left > 0:
        k = min(rows_per_strip, left)
        sizes.append(k)
        left -= k
    if tail:
        sizes.append(tail)
    return sizes


def plan_windows(H: int, W: int, patch: int = PATCH, overlap: int = OVERLAP, merge: bool = True,
                 rows_per_strip: int = 2, first_strip_rows: Optional[int] = None, last_strip_rows: Optional[int] = None) -> List[Window]:
    """All windows covering the raster the way get_patch_indices does (PopulationDataset.py:294-316).
    first_strip_rows / last_strip_rows: tile-rows of the first / last merged strip (default rows_per_strip), see strip_sizes."""
    if H < patch or W < patch:
        raise ValueError(f"raster {H}x{W} is smaller than the inference patch {patch}")
    xs, ys = grid_origins(H, patch, overlap), grid_origins(W, patch, overlap)
    max_x, max_y = H - patch, W - patch
    stride = patch - 2 * overlap
    wins: List[Window] = []
    if not merge:
        for i, x in enumerate(xs):
            for y in ys:
                wins.append(Window(x, y, patch, patch, i, 1))
            wins.append(Window(x, max_y, patch, patch, i, 1))                    # right column
        for y in ys:
            wins.append(Window(max_x, y, patch, patch, -1, 1))                   # bottom row
        wins.append(Window(max_x, max_y, patch, patch, -1, 1))                   # corner
        return wins
    width = (ys[-1] + patch) if ys else 0
    starts, i0 = [], 0
    for k in strip_sizes(len(xs), rows_per_strip, first_strip_rows, last_strip_rows):
        starts.append((i0, k))
        i0 += k
    for i0, k in starts:
        height = stride * (k - 1) + patch
        if ys:
            wins.append(Window(xs[i0], 0, height, width, i0, k * len(ys)))
        wins.append(Window(xs[i0], max_y, height, patch, i0, k))                 # right column of these rows
    if ys:
        wins.append(Window(max_x, 0, patch, width, -1, len(ys)))                 # bottom row
    wins.append(Window(max_x, max_y, patch, patch, -1, 1))                       # corner
    return wins


def shard_windows(wins: Sequence[Window], n_tile_rows: int, rank: int, world: int,
                  rows_per_strip: int = 2) -> List[Window]:
    """Contiguous blocks of strips per rank (strip = rows_per_strip tile-rows); bottom edge row -> last owner."""
    strips = sorted({w.tile_row for w in wins if w.tile_row >= 0})
    if not strips:
        return list(wins) if rank == 0 else []
    per = [len(strips) // world + (1 if r < len(strips) % world else 0) for r in range(world)]
    start = sum(per[:rank])
    mine = set(strips[start:start + per[rank]])
    last_owner = max(r for r in range(world) if per[r] > 0)
    return [w for w in wins if (w.tile_row in mine) or (w.tile_row < 0 and rank == last_owner)]


def plan_balanced_shards(H: int, W: int, world: int, patch: int = PATCH, overlap: int = OVERLAP, rows_per_strip: int = 2,
                         unit: int = 256, first_strip_rows: Optional[int] = None) -> List[List[Window]]:
    """Per-rank windows with the main grid split at ``unit``-row granularity instead of whole tile-rows.

    ``shard_windows`` hands out whole strips, so with 27 tile-rows on 8 GPUs (Uganda) the busiest rank computes 4 tile-rows
    where 3.4 would do.  Here the rows the main grid writes, [overlap, overlap + stride * n_tile_rows), are cut into units of
    ``unit`` rows (stride % unit == 0, unit % 4 == 0: every window origin stays on the reference grid's pool phase, and with
    unit a multiple of the kernels' 64-row CTA tiles also on their tiling), and each rank gets a contiguous run of units
    chosen so that the *input rows it computes* (units + 2*overlap of halo per window + the 2048-row bottom edge windows on
    the last rank) are as equal as possible.  The last rank always owns the rows the bottom edge tiles overlap, so both
    contributors of a doubly covered pixel stay on one GPU.  Written pixels, visit counts and values are those of the
    reference tile grid (tests/test_country_host.py); ``Window.ntiles`` is 0 for these windows (they are not unions of tiles).
    """
    if H < patch or W < patch:
        raise ValueError(f"raster {H}x{W} is smaller than the inference patch {patch}")
    stride = patch - 2 * overlap
    if not can_merge(patch, overlap) or stride % unit or unit % 4:
        raise ValueError(f"balanced sharding needs mergeable tiles and stride {stride} % unit {unit} == 0 (unit % 4 == 0)")
    xs, ys = grid_origins(H, patch, overlap), grid_origins(W, patch, overlap)
    max_x, max_y = H - patch, W - patch
    width = (ys[-1] + patch) if ys else 0
    edge = ([Window(max_x, 0, patch, width, -1, len(ys))] if ys else []) + [Window(max_x, max_y, patch, patch, -1, 1)]
    if not xs:
        return [edge if r == 0 else [] for r in range(world)]
    n_units = len(xs) * stride // unit
    max_u = max(1, rows_per_strip * stride // unit)               # units per window (bounds the activation workspace)
    halo = 2.0 * overlap / unit                                    # recomputed rows per window, in units
    need_last = max(0, -(-(stride * len(xs) - max_x) // unit))     # rows shared with the bottom edge tiles -> last rank
    units = [0] * world
    units[-1] = min(n_units, need_last)

    def cost(r: int, u: int) -> float:
        c = u + halo * (-(-u // max_u)) if u else 0.0
        return c + (patch / unit if r == world - 1 else 0.0)

    for _ in range(n_units - units[-1]):                           # hand out unit by unit to the least loaded rank
        r = min(range(world), key=lambda q: (cost(q, units[q] + 1), q))
        units[r] += 1
    shards: List[List[Window]] = []
    a = overlap
    for r in range(world):
        wins: List[Window] = []
        end = a + units[r] * unit
        first = True
        while a < end:
            k = max_u
            if first and r == 0 and first_strip_rows:
                k = max(1, first_strip_rows * stride // unit)
            b = min(end, a + k * unit)
            y0, h = a - overlap, b - a + 2 * overlap
            if ys:
                wins.append(Window(y0, 0, h, width, y0 // stride, 0))
            wins.append(Window(y0, max_y, h, patch, y0 // stride, 0))          # right column of these rows
            a, first = b, False
        if r == world - 1:
            wins += edge
        shards.append(wins)
    return shards


def owned_rows(wins: Sequence[Window], H: int, overlap: int = OVERLAP) -> Tuple[int, int]:
    """Raster rows [lo, hi) this rank writes (centre rows of its windows)."""
    if not wins:
        return 0, 0
    return min(w.y0 + overlap for w in wins), max(w.y0 + w.h - overlap for w in wins)


def input_rows(wins: Sequence[Window]) -> Tuple[int, int]:
    if not wins:
        return 0, 0
    return min(w.y0 for w in wins), max(w.y0 + w.h for w in wins)


def allreduce_sums(sums: torch.Tensor, group=None) -> torch.Tensor:
    """The path's only collective: SUM of the R partial region sums (fp64) over all ranks."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


class RawRaster:
    """A raster frame in its on-disk dtypes: Sentinel-2 [4, rows, W] uint16 (or float32) and Sentinel-1
    [2, rows, W] float32, both on the device or both in pinned host memory.  ``s2_plane_map`` gives the plane of each
    output channel R,G,B,NIR (GeoTIFF band order B02,B03,B04,B08 -> ops.S2_FILE_TO_RGBN).  CountryEngine.run uploads
    16 B/pixel instead of 24 and converts + normalises on the device (csrc/ingest.cu), bit-identical to the reference's
    .astype(float32) + apply_normalize (data/PopulationDataset.py:594-604, utils/utils.py:105-127).

    PRECONDITION — no NaNs.  The reference fills NaNs per tile before the model sees them (interpolate_nan, fall back to the other
    orbit, raise above 5 %: data/PopulationDataset.py:477-499, 526-551); that loader logic is out of this path's scope.
    A NaN in S1 (swath edges) would spread through the 23-px receptive field of every window that contains it and turn the fp64
    census sums of whole regions — and the all-reduced totals — into NaN.  Fill before constructing a RawRaster; ``assert_finite()``
    checks it (one pass over S1, not on the hot path)."""

    def __init__(self, s2: torch.Tensor, s1: torch.Tensor, s2_plane_map: int = ops.S2_FILE_TO_RGBN, stats: Optional[dict] = None):
        if s2.dim() != 3 or s1.dim() != 3 or s2.shape[0] != 4 or s1.shape[0] != 2 or s2.shape[1:] != s1.shape[1:]:
            raise ValueError("RawRaster needs S2 [4,rows,W] and S1 [2,rows,W] of the same extent")
        if s2.dtype not in (torch.uint16, torch.float32) or s1.dtype != torch.float32:
            raise ValueError("RawRaster: S2 must be uint16 or float32, S1 float32")
        if s2.is_cuda != s1.is_cuda:
            raise ValueError("RawRaster: S2 and S1 must live on the same side (device, or pinned host)")
        self.s2, self.s1, self.s2_plane_map, self.stats = s2, s1, s2_plane_map, stats
        self.is_cuda = s2.is_cuda
        self.shape = (6,) + tuple(s2.shape[1:])

    def is_pinned(self) -> bool:
        return self.s2.is_pinned() and self.s1.is_pinned()

    def assert_finite(self) -> "RawRaster":
        """Raise ValueError if S1 (or a float32 S2) holds NaN / Inf (see the class docstring); returns self."""
        for name, t in (("S1", self.s1), ("S2", self.s2)):
            if t.is_floating_point() and not bool(torch.isfinite(t).all()):
                raise ValueError(f"RawRaster: {name} contains NaN / Inf — fill them first (reference: PopulationDataset.interpolate_nan)")
        return self


class CountryEngine:
    """Tiled inference of one raster frame with an ensemble of POPCORN members on the current CUDA device."""

    def __init__(self, models, H: int, W: int, patch: int = PATCH, overlap: int = OVERLAP, merge: bool = True,
                 rows_per_strip: int = 2, rank: int = 0, world: int = 1, want_scale: bool = True,
                 want_std: bool = True, first_strip_rows: Optional[int] = None, balance: bool = False,
                 balance_unit: int = 256, upload_once: bool = False, last_strip_rows: Optional[int] = None):
        self.models = list(models) if isinstance(models, (list, tuple)) else [models]
        self.H, self.W, self.patch, self.overlap = H, W, patch, overlap
        self.rank, self.world = rank, world
        self.want_scale, self.want_std = want_scale, want_std
        self.upload_once = upload_once      # opt-in: host RawRaster rows cross PCIe once (see _run_resident)
        merge = merge and can_merge(patch, overlap)      # otherwise fall back to the reference tile grid
        self.merged = merge
        all_w = plan_windows(H, W, patch, overlap, merge, rows_per_strip if merge else 1, first_strip_rows if merge else None,
                             last_strip_rows if merge else None)
        n_rows = len(grid_origins(H, patch, overlap))
        if balance and world > 1:     # opt-in: unit-row granularity instead of whole strips (plan_balanced_shards)
            if not merge:
                raise ValueError("balance=True needs merged windows")
            self.windows = plan_balanced_shards(H, W, world, patch, overlap, rows_per_strip, balance_unit,
                                                first_strip_rows)[rank]
        else:
            self.windows = shard_windows(all_w, n_rows, rank, world, rows_per_strip if merge else 1) if world > 1 else all_w
        self.out_rows = owned_rows(self.windows, H, overlap)
        self.in_rows = input_rows(self.windows)
        self._maps = None
        self._copy_stream = None
        self._d2h_stream = None
        # ensemble members share the builtup pass when their building_extractor weights are identical
        # (they never receive gradients: model/popcorn.py:112-114) — SURVEY.md §8f N2
        self._bext_shared = len(self.models) > 1 and all(
            torch.equal(self.models[0]._dda_pack("building_extractor"), m._dda_pack("building_extractor"))
            for m in self.models[1:])

    # ------------------------------------------------------------------------------------------
    def alloc_maps(self, device):
        lo, hi = self.out_rows
        n = max(hi - lo, 0)
        z = lambda dt=torch.float32: torch.zeros(n, self.W, dtype=dt, device=device)
        self._maps = [z(), z() if self.want_std else None, z() if self.want_scale else None,
                      z() if (self.want_scale and self.want_std) else None, z(torch.int16)]
        return self._maps

    def _forward_window(self, x: torch.Tensor, win: Window):
        """x [1,6,h,w] device view.  Runs every member; accumulates the centre into the maps."""
        ov = self.overlap
        lo, _ = self.out_rows
        builtup = None
        for mi, m in enumerate(self.models):
            if builtup is None or not self._bext_shared:
                builtup = ops.dda_forward(m._dda_pack("building_extractor"), x, m.p2d, ops.PC_DDA_BUILTUP)
            feats = ops.dda_forward(m._dda_pack("unetmodel"), x, (0, 0, 0, 0), ops.PC_DDA_FEATURES)
            bu = builtup if m.occupancymodel else None
            tc = popcorn_mod.USE_TENSOR_CORE_HEAD
            dens, scale = ops.head_dense_forward(m._head_pack(tc=tc), feats, bu, None, None, None,
                                                 want_scale=self.want_scale and m.occupancymodel, tc=tc)
            ops.accumulate_tile(dens[0], None if scale is None else scale[0], (ov, win.h - ov), (ov, win.w - ov),
                                self._maps, win.y0 - lo, win.x0)
            del feats, dens, scale

    def _rows_final_after(self, k: int) -> int:
        """Owned-map rows [0, n) that no window after windows[k] writes again (windows are ordered by strip)."""
        lo, hi = self.out_rows
        later = self.windows[k + 1:]
        if not later:
            return hi - lo
        return max(0, min(w.y0 + self.overlap for w in later) - lo)

    def _ship_rows(self, r0: int, r1: int, map_out: torch.Tensor, dev):
        """Finalise map rows [r0, r1) and start their device->host copy on the download stream."""
        if r1 <= r0:
            return
        ops.finalize_map(self._maps, rows=(r0, r1))
        if self._d2h_stream is None:
            self._d2h_stream = torch.cuda.Stream(device=dev)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        self._d2h_stream.wait_event(ev)
        ops.copy_d2h(map_out[r0:r1], self._maps[0][r0:r1], self._d2h_stream.cuda_stream)
        if self._maps[0].is_cuda:
            self._maps[0].record_stream(self._d2h_stream)      # the next run re-allocates the maps: this block is not reused before the copy is done

    def run(self, raster: torch.Tensor, ids: Optional[torch.Tensor], R: int, row_offset: int = 0,
            group=None, finalize: bool = True, map_out: Optional[torch.Tensor] = None, reduce: bool = True):
        """raster: [6, rows, W] fp32 normalised, holding raster rows [row_offset, row_offset+rows) — a CUDA tensor, or a
        pinned host tensor (then windows are streamed H2D on a copy stream, overlapped with compute) — or a RawRaster
        (uint16 S2 + float32 S1 in on-disk form, device or pinned host: converted and normalised on the device).
        ids: int32 [out_hi-out_lo, W] CUDA id raster for this rank's owned rows (or None).
        map_out: optional pinned host tensor [owned rows, W]: finished strips are finalised and copied out while later
        strips still compute (replaces the per-tile .cpu() of run_eval.py:127-135); the copies are asynchronous — call
        engine.wait_download() (or torch.cuda.synchronize()) before reading map_out.
        reduce=False skips the all-reduce (the returned sums are this rank's partial sums).
        Returns dict(map, std, scale_map, scale_std, count, sums[R] float64 all-reduced)."""
        if map_out is not None:
            if not finalize:
                raise ValueError("map_out needs finalize=True")
            if map_out.is_cuda or (map_out.numel() and not map_out.is_pinned()) or tuple(map_out.shape) != (self.out_rows[1] - self.out_rows[0], self.W):
                raise ValueError("map_out must be a pinned host tensor of the owned map shape")
        # a previous run's download may still be in flight: its map block is protected by record_stream (_ship_rows), and copies into
        # the same host buffer stay ordered on the download stream — so consecutive runs pipeline (prefetch() + run() + run() ...)
        self._map_out, self._shipped = map_out, 0
        dev = torch.device("cuda", torch.cuda.current_device())
        maps = self.alloc_maps(dev)
        raw = isinstance(raster, RawRaster)
        if not self.windows:
            pass                                  # this rank owns no rows (more ranks than strips): only the all-reduce below
        elif raster.is_cuda and not raw:
            for k, win in enumerate(self.windows):
                x = raster[None, :, win.y0 - row_offset: win.y0 - row_offset + win.h, win.x0: win.x0 + win.w]
                self._forward_window(x, win)
                self._after_window(k, dev)
        elif raster.is_cuda:
            xbuf = None
            for k, win in enumerate(self.windows):
                r0 = win.y0 - row_offset
                if xbuf is None or xbuf.numel() < 6 * win.h * win.w:
                    xbuf = torch.empty(6 * max(w.h for w in self.windows) * max(w.w for w in self.windows),
                                       dtype=torch.float32, device=dev)
                x = xbuf[: 6 * win.h * win.w].view(6, win.h, win.w)
                ops.ingest_normalize(raster.s2[:, r0: r0 + win.h, win.x0: win.x0 + win.w],
                                     raster.s1[:, r0: r0 + win.h, win.x0: win.x0 + win.w], x, raster.s2_plane_map, raster.stats)
                self._forward_window(x[None], win)
                self._after_window(k, dev)
        elif raw and self.upload_once:
            self._run_resident(raster, row_offset, dev)
        else:
            self._run_streamed(raster, row_offset, dev)
        if finalize:
            if map_out is not None:
                self._ship_rows(self._shipped, self.out_rows[1] - self.out_rows[0], map_out, dev)
            else:
                ops.finalize_map(maps)
        sums = torch.zeros(max(R, 1), dtype=torch.float64, device=dev)
        if ids is not None and maps[0].numel():
            ops.region_sum(maps[0], ids, R, sums)
        if reduce:
            allreduce_sums(sums, group)
        return {"map": maps[0], "std": maps[1], "scale_map": maps[2], "scale_std": maps[3], "count": maps[4],
                "sums": sums, "rows": self.out_rows}

    def _after_window(self, k: int, dev):
        if self._map_out is None:
            return
        n = self._rows_final_after(k)
        if n > self._shipped:
            self._ship_rows(self._shipped, n, self._map_out, dev)
            self._shipped = n

    def wait_download(self):
        """Block until the map rows shipped through ``map_out`` have landed in host memory."""
        if self._d2h_stream is not None:
            self._d2h_stream.synchronize()

    def _run_streamed(self, raster: torch.Tensor, row_offset: int, dev):
        """Host raster -> device, one window ahead of the compute (double buffer, copy stream + events)."""
        if not raster.is_pinned():
            raise RuntimeError("host rasters must be pinned (torch.Tensor.pin_memory) for the streamed path")
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        cs, main = self._copy_stream, torch.cuda.current_stream(dev)
        wins = self.windows
        if not wins:
            return
        # ring of NB device buffers, uploads run NB-1 windows ahead: a big strip's upload must overlap the previous
        # big strip's compute, not just the small right-column window that sits between them
        NB = 3
        mh, mw = max(w.h for w in wins), max(w.w for w in wins)
        raw = isinstance(raster, RawRaster)
        if raw:   # raw bands travel in their on-disk dtypes (16 B/px); one fp32 window is produced on the device per forward
            bufs = [(torch.empty(4 * mh * mw, dtype=raster.s2.dtype, device=dev), torch.empty(2 * mh * mw, dtype=torch.float32, device=dev))
                    for _ in range(NB)]
            xnorm = torch.empty(6 * mh * mw, dtype=torch.float32, device=dev)
        else:
            bufs = [torch.empty(6 * mh * mw, dtype=torch.float32, device=dev) for _ in range(NB)]
        ready = [torch.cuda.Event() for _ in range(NB)]
        free = [torch.cuda.Event() for _ in range(NB)]
        views = {}

        def upload(k):
            w = wins[k]
            r0 = w.y0 - row_offset
            with torch.cuda.stream(cs):
                cs.wait_event(free[k % NB])
                if raw:
                    b2 = bufs[k % NB][0][: 4 * w.h * w.w].view(4, w.h, w.w)
                    b1 = bufs[k % NB][1][: 2 * w.h * w.w].view(2, w.h, w.w)
                    ops.copy_window_h2d(b2, raster.s2[:, r0: r0 + w.h, w.x0: w.x0 + w.w], cs.cuda_stream)
                    ops.copy_window_h2d(b1, raster.s1[:, r0: r0 + w.h, w.x0: w.x0 + w.w], cs.cuda_stream)
                    b = (b2, b1)
                else:
                    b = bufs[k % NB][: 6 * w.h * w.w].view(6, w.h, w.w)
                    ops.copy_window_h2d(b, raster[:, r0: r0 + w.h, w.x0: w.x0 + w.w], cs.cuda_stream)
                ready[k % NB].record(cs)
            views[k] = b

        for e in free:
            e.record(main)
        for k in range(min(NB - 1, len(wins))):
            upload(k)
        for k, win in enumerate(wins):
            if k + NB - 1 < len(wins):
                upload(k + NB - 1)
            main.wait_event(ready[k % NB])
            b = views.pop(k)
            if raw:
                x = xnorm[: 6 * win.h * win.w].view(6, win.h, win.w)
                ops.ingest_normalize(b[0], b[1], x, raster.s2_plane_map, raster.stats)
                free[k % NB].record(main)          # the raw buffers are consumed by the ingest kernel
                self._forward_window(x[None], win)
            else:
                self._forward_window(b[None], win)
                free[k % NB].record(main)
            self._after_window(k, dev)
        px = sum(w.h * w.w for w in wins)
        self.h2d_bytes = px * (4 * raster.s2.element_size() + 2 * 4) if raw else px * 6 * 4

    def _slab_set(self, k: int, raster: "RawRaster", dev):
        """Device slabs (on-disk dtypes, this rank's input rows) of buffer set k, allocated once and kept across runs."""
        i0, i1 = self.in_rows
        rows, W = i1 - i0, self.W
        if not hasattr(self, "_slabs"):
            self._slabs = [None, None]
            self._slab_free = [None, None]      # event: the last ingest kernel that read set k has been queued
        cur = self._slabs[k]
        if cur is None or cur[0].dtype != raster.s2.dtype or cur[0].shape != (4, rows, W):
            self._slabs[k] = (torch.empty(4, rows, W, dtype=raster.s2.dtype, device=dev), torch.empty(2, rows, W, dtype=torch.float32, device=dev))
            # fresh memory from the caching allocator may still be read by work queued on the current stream: the first upload waits for it
            self._slab_free[k] = torch.cuda.Event()
            self._slab_free[k].record(torch.cuda.current_stream(dev))
        return self._slabs[k]

    def _upload_slabs(self, k: int, raster: "RawRaster", row_offset: int, dev, chunk_rows: int):
        """Queue the upload of this rank's raw rows into slab set k on the copy stream, in row order; -> one 'landed' event per chunk."""
        if not raster.is_pinned():
            raise RuntimeError("host rasters must be pinned (torch.Tensor.pin_memory) for the streamed path")
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        cs, main = self._copy_stream, torch.cuda.current_stream(dev)
        i0, i1 = self.in_rows
        rows, W = i1 - i0, self.W
        if i0 - row_offset < 0 or i1 - row_offset > raster.shape[1] or raster.shape[2] != W:
            raise ValueError("raster does not hold this rank's input rows")
        d2, d1 = self._slab_set(k, raster, dev)
        landed = []
        with torch.cuda.stream(cs):
            # The upload waits ONLY for the last ingest kernel that read this slab set (or, for fresh slabs, for what was queued when they
            # were allocated) — not for the run in flight, which reads the other set: prefetch() called right after run() must overlap it.
            # (Waiting for "everything queued so far" serialised the next raster's upload behind the current run: 20 ms per step.)
            if self._slab_free[k] is not None:
                cs.wait_event(self._slab_free[k])
            for a in range(0, rows, chunk_rows):
                b = min(rows, a + chunk_rows)
                h0 = i0 - row_offset + a
                ops.copy_window_h2d(d2[:, a:b], raster.s2[:, h0: h0 + b - a], cs.cuda_stream)
                ops.copy_window_h2d(d1[:, a:b], raster.s1[:, h0: h0 + b - a], cs.cuda_stream)
                ev = torch.cuda.Event()
                ev.record(cs)
                landed.append(ev)
        return landed

    def prefetch(self, raster: "RawRaster", row_offset: int = 0, chunk_rows: int = 512):
        """Start uploading the NEXT raster (pinned host RawRaster, upload_once engines) while the current run still computes: the
        copies go to the other slab set on the copy stream; the next ``run(raster, ...)`` with the same object picks them up.
        A stream of rasters (seasonal frames, successive countries) then pays the upload bubble once, not per raster."""
        if not (isinstance(raster, RawRaster) and not raster.is_cuda and self.upload_once):
            raise ValueError("prefetch() needs a pinned host RawRaster and an engine built with upload_once=True")
        if not self.windows:
            return
        dev = torch.device("cuda", torch.cuda.current_device())
        k = 1 - getattr(self, "_slab_cur", 1)
        landed = self._upload_slabs(k, raster, row_offset, dev, chunk_rows)
        self._prefetched = (raster, row_offset, k, landed, chunk_rows)

    def _run_resident(self, raster: "RawRaster", row_offset: int, dev, chunk_rows: int = 512):
        """Host RawRaster -> device, every input row ONCE: the rank's rows are copied in row chunks (copy stream, in row
        order) into device slabs that keep the on-disk dtypes (16 B/px: Uganda on 8 GPUs = 5 GB per rank), and each window
        is converted + normalised from the slab as soon as the chunks it needs have landed.  The window-by-window upload
        of _run_streamed sends the overlapping halos and the right-column windows again (1.18-1.20x the unique rows),
        which is what bounds the end-to-end rate once 8 GPUs share the host's PCIe complex.  Two slab sets alternate, so
        ``prefetch()`` can upload the next raster during this run."""
        main = torch.cuda.current_stream(dev)
        wins = self.windows
        i0, i1 = self.in_rows
        rows, W = i1 - i0, self.W
        pf = getattr(self, "_prefetched", None)
        if pf is not None and pf[0] is raster and pf[1] == row_offset:
            _, _, k, landed, chunk_rows = pf
        else:
            k = 1 - getattr(self, "_slab_cur", 1)
            landed = self._upload_slabs(k, raster, row_offset, dev, chunk_rows)
        self._prefetched = None
        self._slab_cur = k
        d2, d1 = self._slabs[k]
        mh, mw = max(w.h for w in wins), max(w.w for w in wins)
        if getattr(self, "_xnorm", None) is None or self._xnorm.numel() < 6 * mh * mw or self._xnorm.device != dev:
            self._xnorm = torch.empty(6 * mh * mw, dtype=torch.float32, device=dev)
        xnorm = self._xnorm
        for k_w, win in enumerate(wins):
            r0 = win.y0 - i0
            main.wait_event(landed[(r0 + win.h - 1) // chunk_rows])        # copies are in row order on one stream
            x = xnorm[: 6 * win.h * win.w].view(6, win.h, win.w)
            ops.ingest_normalize(d2[:, r0: r0 + win.h, win.x0: win.x0 + win.w], d1[:, r0: r0 + win.h, win.x0: win.x0 + win.w],
                                 x, raster.s2_plane_map, raster.stats)
            self._forward_window(x[None], win)
            self._after_window(k_w, dev)
        ev = torch.cuda.Event()
        ev.record(main)
        self._slab_free[k] = ev                    # a later upload into this set waits for the last ingest kernel
        self.h2d_bytes = rows * W * (4 * raster.s2.element_size() + 2 * 4)


def adjust_map_to_census(map_: torch.Tensor, ids: torch.Tensor, sums: torch.Tensor, census_pop: torch.Tensor):
    """Dasymetric rescale (PopulationDataset.py:842-850) as one gather-multiply pass: region r is scaled by
    census_pop[r] / sums[r] (left unchanged where the predicted sum is 0).  ids int32, id 0..R-1."""
    s = sums.to(torch.float32)
    factor = torch.where(s == 0, torch.ones_like(s), census_pop.to(s.device, torch.float32) / s)
    return ops.region_scale_(map_, ids, factor.contiguous())
