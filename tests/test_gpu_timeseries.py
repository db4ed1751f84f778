"""GPU: multi-temporal inference (BASELINE configs[4]) against the notebook's statistics computed with the oracle:
per seasonal frame every ensemble member, ensemble mean / unbiased std, frame totals, season average
(time_series_inference.ipynb nb-lines 191-245)."""
import pytest
import torch

from popcorn_b200 import country as ct
from popcorn_b200 import timeseries as ts
from oracle import popcorn_oracle as po
from util import TOL_PIXEL, TOL_REGION, build_model, golden_state_dict, max_rel

pytestmark = pytest.mark.gpu


def _members(n):
    """Ensemble members = the golden state_dict with differently seeded heads (the builtup extractor is shared)."""
    sds = []
    for i in range(n):
        sd = {k: v.clone() for k, v in golden_state_dict().items()}
        g = torch.Generator().manual_seed(100 + i)
        for k in sd:
            if k.startswith("head.") and k.endswith("weight"):
                sd[k] = sd[k] + 0.05 * torch.randn(sd[k].shape, generator=g)
        sds.append(sd)
    return sds


def test_whole_raster_frames_match_the_notebook_statistics():
    """One forward per member on the whole (odd-sized) raster, like the notebook: mean, torch.std, totals, season mean."""
    H, W, T = 75, 101, 3
    sds = _members(2)
    models = [build_model(sd).eval() for sd in sds]
    frames = [po.synthetic_input(H, W, seed=40 + t) for t in range(T)]
    means, stds = [], []
    for x in frames:
        m, s = ts.whole_raster_frame(models, x.cuda())
        ref = torch.stack([po.forward(sd, {"input": x.clone()}, padding=False)["popdensemap"][0] for sd in sds])
        assert max_rel(m, ref.mean(0)) < TOL_PIXEL
        assert max_rel(s, ref.std(0), floor_frac=1e-2) < 5e-2      # std of two nearly equal maps: cancellation amplifies fp32 noise
        assert abs(float(m.sum()) - float(ref.mean(0).sum())) < TOL_REGION * float(ref.mean(0).sum())
        means.append(ref.mean(0))
    season_ref = torch.stack(means).mean(0)
    season = torch.stack([ts.whole_raster_frame(models, x.cuda())[0] for x in frames]).mean(0)
    assert max_rel(season, season_ref) < TOL_PIXEL


def test_tiled_time_series_engine_matches_per_frame_engine_runs():
    H, W, ps, ov, T, R = 520, 456, 128, 32, 4, 20
    models = [build_model(sd).eval() for sd in _members(2)]
    ids = po.synthetic_regions(H, W, R).cuda()
    frames = []
    for t in range(T):
        s2, s1 = po.synthetic_raw(H, W, seed=70 + t)
        frames.append(ct.RawRaster(s2.cuda(), s1.cuda()) if t % 2 == 0 else po.read_and_normalize(s2, s1)[0].cuda())
    eng = ts.TimeSeriesEngine(models, H, W, patch=ps, overlap=ov, merge=True, rows_per_strip=2)
    lo, hi = eng.out_rows
    with torch.no_grad():
        out = eng.run(frames, ids[lo:hi].contiguous(), R + 1, keep_frames=True)
        single = ct.CountryEngine(models, H, W, ps, ov, merge=True, rows_per_strip=2)
        maps, stds, sums = [], [], []
        for fr in frames:
            o = single.run(fr, ids[lo:hi].contiguous(), R + 1)
            maps.append(o["map"].clone()); stds.append(o["std"].clone()); sums.append(o["sums"].clone())
    for a, b in zip(out["frame_maps"], maps):
        assert torch.equal(a, b)
    assert torch.allclose(out["season_map"], torch.stack(maps).mean(0), rtol=1e-6, atol=1e-9)
    assert torch.allclose(out["season_std"], torch.stack(stds).mean(0), rtol=1e-5, atol=1e-9, equal_nan=True)
    assert torch.allclose(out["totals"], torch.stack([m.sum(dtype=torch.float64) for m in maps]), rtol=1e-9)
    assert torch.allclose(out["sums"], torch.stack(sums), rtol=1e-9)
    assert abs(float(out["season_total"]) - float(out["totals"].mean())) < 1e-6 * float(out["totals"].mean())
    # census sums of a frame == the frame total over the written (non-frame) region with ids > 0 ... and ensemble std is finite
    assert bool(torch.isfinite(out["season_std"][out["season_std"] == out["season_std"]]).all())
