"""Multi-temporal (seasonal) inference: BASELINE.json configs[4], modelled on time_series_inference.ipynb.

The notebook runs, per year and per season (spring, summer, autumn, winter): every ensemble member on the frame
(nb-lines 191-196: ``thismodel(sample, padding=False)["popdensemap"]``), the ensemble mean and unbiased std
(nb-lines 198-202), the frame total ``popdense_mean.sum()`` (nb-line 209), and then the season average of the means and
of the stds and its total (nb-lines 235-245).  Here every frame goes through the tiled CountryEngine (so rasters far
beyond one forward's memory work, rows sharded over the ranks), the per-frame maps never leave the device, the season
accumulation is an in-place axpy on the device.  On several GPUs the series is partitioned on two axes, frames x row strips
(``plan_frames``); the collectives are the engine's all-reduce of the R region sums per frame, one all-reduce of the season
maps across the frame groups, and one all-reduce of the T+1 totals at the end.

For rasters small enough for ONE forward per member (the notebook's own case) ``whole_raster_frame`` reproduces the
notebook call literally (reflect padding to a multiple of 64 inside POPCORN.forward, no tiling frame).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .country import CountryEngine, RawRaster, allreduce_sums


def whole_raster_frame(models: Sequence, x: torch.Tensor):
    """x [1,6,H,W] normalised CUDA tensor -> (ensemble mean [H,W], unbiased ensemble std [H,W]) as nb-lines 191-202."""
    maps = []
    with torch.no_grad():
        for m in models:
            maps.append(m({"input": x}, padding=False)["popdensemap"][0])
    stack = torch.stack(maps)
    mean = stack.mean(dim=0)
    std = stack.std(dim=0) if len(maps) > 1 else torch.full_like(mean, float("nan"))   # torch.std of one sample is NaN
    return mean, std


def plan_frames(T: int, world: int):
    """Partition of a T-frame series over `world` ranks on two independent axes (frames x row strips).
    F = the largest divisor of `world` that is <= T frame groups; each group holds S = world // F ranks that shard the ROWS of
    the group's frames.  Rank r -> (frame group r // S, row shard r % S); frame t belongs to group t % F.
    Returns (F, S, [(frame_group, row_shard, [frames]) for every rank]).  8 ranks x 4 frames -> 4 groups of 2 row shards: every
    rank works, where rows-only sharding leaves half of them without a strip (4 strips in a Switzerland-shaped raster)."""
    if T < 1 or world < 1:
        raise ValueError("need at least one frame and one rank")
    F = max(d for d in range(1, world + 1) if world % d == 0 and d <= T)
    S = world // F
    return F, S, [(r // S, r % S, [t for t in range(T) if t % F == r // S]) for r in range(world)]


class TimeSeriesEngine:
    """Seasonal frames of one raster -> per-frame maps / totals / census sums and their season average.

    Multi-GPU: frames x row strips (plan_frames).  Collectives: the engine's all-reduce of the R census sums per frame inside a frame
    group; ONE all-reduce of the season mean / std maps across the frame groups that hold the same rows (the season average of
    nb-line 238 is a real exchange between frames: rows x W floats over NVLink); one all-reduce of the T+1 totals."""

    def __init__(self, models, H: int, W: int, rank: int = 0, world: int = 1, frames: Optional[int] = None, **engine_kw):
        self.H, self.W, self.rank, self.world = H, W, rank, world
        self.T = frames
        self.F, self.S = 1, world
        self.frame_group, self.row_shard = 0, rank
        self.my_frames = None if frames is None else list(range(frames))
        self._row_group = self._frame_axis_group = None
        if frames is not None and world > 1:
            self.F, self.S, plan = plan_frames(frames, world)
            self.frame_group, self.row_shard, self.my_frames = plan[rank]
            if self.F > 1:
                import torch.distributed as dist
                if not (dist.is_available() and dist.is_initialized()):
                    raise RuntimeError("TimeSeriesEngine(frames=..., world>1) needs an initialised torch.distributed process group")
                # new_group is collective: every rank creates every group, in the same order
                for g in range(self.F):                       # ranks of one frame group: shard the rows of its frames
                    grp = dist.new_group([g * self.S + k for k in range(self.S)])
                    if g == self.frame_group:
                        self._row_group = grp
                for k in range(self.S):                       # ranks that hold the same rows in different frame groups
                    grp = dist.new_group([g * self.S + k for g in range(self.F)])
                    if k == self.row_shard:
                        self._frame_axis_group = grp
        self.engine = CountryEngine(models, H, W, rank=self.row_shard, world=self.S, **engine_kw)

    @property
    def out_rows(self):
        return self.engine.out_rows

    @property
    def in_rows(self):
        return self.engine.in_rows

    def describe(self) -> str:
        return f"{self.F} frame group(s) x {self.S} row shard(s)"

    def run(self, frames, ids: Optional[torch.Tensor] = None, R: int = 0, row_offset: int = 0, group=None,
            keep_frames: bool = False):
        """frames: per season a [6, rows, W] normalised tensor (CUDA or pinned host) or a RawRaster holding this rank's input rows —
        a sequence over all T frames (entries of frames this rank does not own may be None) or a dict {frame index: frame}.
        Returns dict(season_map, season_std, totals[T] (per-frame total population, all ranks), season_total, sums[T,R] census sums
        per frame, season_sums[R], frame_maps (if keep_frames; this rank's frames only))."""
        if isinstance(frames, dict):
            T = self.T if self.T is not None else (max(frames) + 1 if frames else 0)
            get = frames.get
        else:
            T = len(frames)
            get = lambda t: frames[t]
        if T == 0:
            raise ValueError("no frames")
        if self.T is not None and T != self.T:
            raise ValueError(f"engine was planned for {self.T} frames, got {T}")
        mine = self.my_frames if self.my_frames is not None else list(range(T))
        dev = torch.device("cuda", torch.cuda.current_device())
        season_map = season_std = None
        totals = torch.zeros(T + 1, dtype=torch.float64, device=dev)
        sums = torch.zeros(T, max(R, 1), dtype=torch.float64, device=dev)
        kept: List[torch.Tensor] = []
        row_group = self._row_group if self.F > 1 else group
        for t in mine:
            fr = get(t)
            if fr is None:
                raise ValueError(f"frame {t} belongs to this rank (frame group {self.frame_group}) but was not supplied")
            out = self.engine.run(fr, ids, R, row_offset=row_offset, group=row_group)
            m, s = out["map"], out["std"]
            if s is not None:
                # a pixel seen once (one member, no tile overlap) has no ensemble std: the engine's map holds run_eval.py's raw sum of
                # squares there (its division is masked by count > 1, :140-154); the notebook's torch.std of one sample is NaN
                s = torch.where(out["count"] == 1, torch.full((), float("nan"), device=s.device), s)     # never-visited frame pixels stay 0
            totals[t] = m.sum(dtype=torch.float64)
            if self.row_shard == 0 or self.F == 1:
                sums[t] = out["sums"]              # all-reduced inside the frame group: one copy per group enters the frame-axis sum
            if season_map is None:
                season_map = m.clone()
                season_std = None if s is None else s.clone()
            else:
                season_map.add_(m)
                if season_std is not None and s is not None:
                    season_std.add_(s)
            if keep_frames:
                kept.append(m.clone())
        if season_map is None:                    # this rank owns no frame (cannot happen with plan_frames: F <= T)
            lo, hi = self.engine.out_rows
            season_map = torch.zeros(max(hi - lo, 0), self.W, device=dev)
            season_std = torch.zeros_like(season_map) if self.engine.want_std else None
        if self.F > 1:
            import torch.distributed as dist
            # the season average couples the frames: sum the partial season maps of the groups that hold the same rows
            if season_map.numel():
                dist.all_reduce(season_map, group=self._frame_axis_group)
                if season_std is not None:
                    dist.all_reduce(season_std, group=self._frame_axis_group)
            if self.row_shard != 0:
                sums.zero_()
            dist.all_reduce(sums)                  # every rank gets every frame's census sums
        season_map.div_(T)                       # torch.stack(...).mean(dim=0), nb-line 238
        if season_std is not None:
            season_std.div_(T)                   # nb-line 239
        if self.frame_group == 0:
            totals[T] = season_map.sum(dtype=torch.float64)      # the season map is replicated over the frame groups: count it once
        allreduce_sums(totals, group)            # rows (and frames) are sharded: totals are partial per rank
        return {"season_map": season_map, "season_std": season_std, "totals": totals[:T], "season_total": totals[T],
                "sums": sums, "season_sums": sums.mean(dim=0), "frame_maps": kept if keep_frames else None,
                "rows": self.engine.out_rows, "frames": mine}
