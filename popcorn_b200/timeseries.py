"""Multi-temporal (seasonal) inference: BASELINE.json configs[4], modelled on time_series_inference.ipynb.

The notebook runs, per year and per season (spring, summer, autumn, winter): every ensemble member on the frame
(nb-lines 191-196: ``thismodel(sample, padding=False)["popdensemap"]``), the ensemble mean and unbiased std
(nb-lines 198-202), the frame total ``popdense_mean.sum()`` (nb-line 209), and then the season average of the means and
of the stds and its total (nb-lines 235-245).  Here every frame goes through the tiled CountryEngine (so rasters far
beyond one forward's memory work, rows sharded over the ranks), the per-frame maps never leave the device, the season
accumulation is an in-place axpy on the device, and the only collectives are the engine's all-reduce of the R region
sums per frame plus one all-reduce of the T+1 totals at the end.

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


class TimeSeriesEngine:
    """Seasonal frames of one raster -> per-frame maps / totals / census sums and their season average."""

    def __init__(self, models, H: int, W: int, rank: int = 0, world: int = 1, **engine_kw):
        self.engine = CountryEngine(models, H, W, rank=rank, world=world, **engine_kw)
        self.H, self.W = H, W

    @property
    def out_rows(self):
        return self.engine.out_rows

    @property
    def in_rows(self):
        return self.engine.in_rows

    def run(self, frames: Sequence, ids: Optional[torch.Tensor] = None, R: int = 0, row_offset: int = 0, group=None,
            keep_frames: bool = False):
        """frames: per season, a [6, rows, W] normalised tensor (CUDA or pinned host) or a RawRaster holding this rank's
        input rows.  Returns dict(season_map, season_std, totals[T] (per-frame total population, all ranks),
        season_total, sums[T,R] census sums per frame, season_sums[R], frame_maps (if keep_frames))."""
        T = len(frames)
        if T == 0:
            raise ValueError("no frames")
        dev = torch.device("cuda", torch.cuda.current_device())
        season_map = season_std = None
        totals = torch.zeros(T + 1, dtype=torch.float64, device=dev)
        sums = torch.zeros(T, max(R, 1), dtype=torch.float64, device=dev)
        kept: List[torch.Tensor] = []
        for t, fr in enumerate(frames):
            out = self.engine.run(fr, ids, R, row_offset=row_offset, group=group)
            m, s = out["map"], out["std"]
            totals[t] = m.sum(dtype=torch.float64)
            sums[t] = out["sums"]
            if season_map is None:
                season_map = m.clone()
                season_std = None if s is None else s.clone()
            else:
                season_map.add_(m)
                if season_std is not None and s is not None:
                    season_std.add_(s)
            if keep_frames:
                kept.append(m.clone())
        season_map.div_(T)                       # torch.stack(...).mean(dim=0), nb-line 238
        if season_std is not None:
            season_std.div_(T)                   # nb-line 239
        totals[T] = season_map.sum(dtype=torch.float64)
        allreduce_sums(totals, group)            # rows are sharded: totals are partial per rank
        return {"season_map": season_map, "season_std": season_std, "totals": totals[:T], "season_total": totals[T],
                "sums": sums, "season_sums": sums.mean(dim=0), "frame_maps": kept if keep_frames else None,
                "rows": self.engine.out_rows}
