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


def plan_windows(H: int, W: int, patch: int = PATCH, overlap: int = OVERLAP, merge: bool = True,
                 rows_per_strip: int = 2) -> List[Window]:
    """All windows covering the raster the way get_patch_indices does (PopulationDataset.py:294-316)."""
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
    for i0 in range(0, len(xs), rows_per_strip):
        k = min(rows_per_strip, len(xs) - i0)
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


class CountryEngine:
    """Tiled inference of one raster frame with an ensemble of POPCORN members on the current CUDA device."""

    def __init__(self, models, H: int, W: int, patch: int = PATCH, overlap: int = OVERLAP, merge: bool = True,
                 rows_per_strip: int = 2, rank: int = 0, world: int = 1, want_scale: bool = True,
                 want_std: bool = True):
        self.models = list(models) if isinstance(models, (list, tuple)) else [models]
        self.H, self.W, self.patch, self.overlap = H, W, patch, overlap
        self.rank, self.world = rank, world
        self.want_scale, self.want_std = want_scale, want_std
        merge = merge and can_merge(patch, overlap)      # otherwise fall back to the reference tile grid
        self.merged = merge
        all_w = plan_windows(H, W, patch, overlap, merge, rows_per_strip if merge else 1)
        n_rows = len(grid_origins(H, patch, overlap))
        self.windows = shard_windows(all_w, n_rows, rank, world, rows_per_strip if merge else 1) if world > 1 else all_w
        self.out_rows = owned_rows(self.windows, H, overlap)
        self.in_rows = input_rows(self.windows)
        self._maps = None
        self._copy_stream = None
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

    def run(self, raster: torch.Tensor, ids: Optional[torch.Tensor], R: int, row_offset: int = 0,
            group=None, finalize: bool = True):
        """raster: [6, rows, W] fp32 holding raster rows [row_offset, row_offset+rows) — a CUDA tensor, or a pinned
        host tensor (then windows are streamed H2D on a copy stream, overlapped with compute).
        ids: int32 [out_hi-out_lo, W] CUDA id raster for this rank's owned rows (or None).
        Returns dict(map, std, scale_map, scale_std, count, sums[R] float64 all-reduced)."""
        dev = torch.device("cuda", torch.cuda.current_device())
        maps = self.alloc_maps(dev)
        if raster.is_cuda:
            for win in self.windows:
                x = raster[None, :, win.y0 - row_offset: win.y0 - row_offset + win.h, win.x0: win.x0 + win.w]
                self._forward_window(x, win)
        else:
            self._run_streamed(raster, row_offset, dev)
        if finalize:
            ops.finalize_map(maps)
        sums = torch.zeros(max(R, 1), dtype=torch.float64, device=dev)
        if ids is not None and maps[0].numel():
            ops.region_sum(maps[0], ids, R, sums)
        allreduce_sums(sums, group)
        return {"map": maps[0], "std": maps[1], "scale_map": maps[2], "scale_std": maps[3], "count": maps[4],
                "sums": sums, "rows": self.out_rows}

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
        bufs = [torch.empty(6 * mh * mw, dtype=torch.float32, device=dev) for _ in range(NB)]
        ready = [torch.cuda.Event() for _ in range(NB)]
        free = [torch.cuda.Event() for _ in range(NB)]
        views = {}

        def upload(k):
            w = wins[k]
            b = bufs[k % NB][: 6 * w.h * w.w].view(6, w.h, w.w)
            with torch.cuda.stream(cs):
                cs.wait_event(free[k % NB])
                ops.copy_window_h2d(b, raster[:, w.y0 - row_offset: w.y0 - row_offset + w.h, w.x0: w.x0 + w.w],
                                    cs.cuda_stream)
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
            self._forward_window(views.pop(k)[None], win)
            free[k % NB].record(main)
        self.h2d_bytes = sum(6 * w.h * w.w * 4 for w in wins)


def adjust_map_to_census(map_: torch.Tensor, ids: torch.Tensor, sums: torch.Tensor, census_pop: torch.Tensor):
    """Dasymetric rescale (PopulationDataset.py:842-850) as one gather-multiply pass: region r is scaled by
    census_pop[r] / sums[r] (left unchanged where the predicted sum is 0).  ids int32, id 0..R-1."""
    s = sums.to(torch.float32)
    factor = torch.where(s == 0, torch.ones_like(s), census_pop.to(s.device, torch.float32) / s)
    return ops.region_scale_(map_, ids, factor.contiguous())
