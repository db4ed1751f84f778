"""GPU: balanced row shards, ranks without rows and the upload-once input path of CountryEngine (written at the end of round 1,
confirmed on a B200 by the round-1 driver run and again in round 2: all five pass)."""
import pytest
import torch

from popcorn_b200 import country as ct
from popcorn_b200 import timeseries as ts
from oracle import popcorn_oracle as po
from util import build_model, golden_state_dict, max_rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    return build_model(golden_state_dict()).eval()


@pytest.mark.parametrize("world", [2, 3])
def test_balanced_row_shards_are_bit_identical_to_the_unsharded_run(model, world):
    """plan_balanced_shards cuts the main grid at unit-row granularity (origins stay on the pool phase and on the kernels'
    64-row CTA tiling): every written pixel must equal the single-GPU result bit for bit, sums add up."""
    H, W, ps, ov = 1220, 300, 192, 32         # stride 128, unit 64: cuts fall inside tile-rows
    raster = po.synthetic_input(H, W, seed=31)[0].cuda()
    ids = po.synthetic_regions(H, W, 20).cuda()
    with torch.no_grad():
        e0 = ct.CountryEngine([model], H, W, ps, ov, merge=True, rows_per_strip=2)
        flo, fhi = e0.out_rows
        full = e0.run(raster, ids[flo:fhi].contiguous(), 21)
        total = torch.zeros_like(full["sums"])
        seen = 0
        for r in range(world):
            eng = ct.CountryEngine([model], H, W, ps, ov, merge=True, rows_per_strip=2, rank=r, world=world, balance=True,
                                   balance_unit=64)
            lo, hi = eng.out_rows
            i0, i1 = eng.in_rows
            o = eng.run(raster[:, i0:i1].contiguous(), ids[lo:hi].contiguous(), 21, row_offset=i0)
            assert torch.equal(o["map"], full["map"][lo - flo:hi - flo])
            assert torch.equal(o["count"], full["count"][lo - flo:hi - flo])
            total += o["sums"]
            seen += hi - lo
    assert seen == fhi - flo
    assert max_rel(total, full["sums"], floor_frac=1.0) < 1e-6


def test_rank_without_rows_runs_and_contributes_zero(model):
    """More ranks than row strips (config 5 on 8 GPUs): a rank that owns no rows must get through run() — empty maps, zero
    partial sums — so that it still joins the all-reduce instead of raising while its peers wait."""
    H, W, ps, ov = 300, 300, 128, 32          # 3 tile-rows -> ranks 3.. of 8 own nothing
    eng = ct.CountryEngine([model], H, W, ps, ov, merge=True, rows_per_strip=1, rank=6, world=8)
    assert eng.windows == [] and eng.out_rows == (0, 0) and eng.in_rows == (0, 0)
    empty = torch.empty(6, 0, W, device="cuda")
    ids = torch.empty(0, W, dtype=torch.int32, device="cuda")
    with torch.no_grad():
        out = eng.run(empty, ids, 5)
        assert out["map"].shape == (0, W) and float(out["sums"].abs().sum()) == 0.0
        host_map = torch.empty(0, W).pin_memory()
        out = eng.run(empty.cpu().pin_memory(), ids, 5, map_out=host_map)
        eng.wait_download()
        tse = ts.TimeSeriesEngine([model], H, W, rank=6, world=8, patch=ps, overlap=ov, merge=True, rows_per_strip=1)
        o = tse.run([empty, empty], None, 0)
    assert float(o["season_total"]) == 0.0 and o["season_map"].shape == (0, W)


@pytest.mark.parametrize("world,rank", [(1, 0), (2, 1)])
def test_upload_once_equals_the_per_window_upload(model, world, rank):
    """CountryEngine(upload_once=True): every raw input row crosses PCIe once into device slabs (chunked, copy stream) and the
    windows are normalised from the slabs — same maps, sums and host download as the window-by-window streamed upload."""
    H, W, ps, ov = 900, 456, 128, 32
    s2_file, s1 = po.synthetic_raw(H, W, seed=35)
    ids = po.synthetic_regions(H, W, 30).cuda()
    outs = []
    for once in (False, True):
        eng = ct.CountryEngine([model], H, W, ps, ov, merge=True, rows_per_strip=2, rank=rank, world=world, upload_once=once)
        lo, hi = eng.out_rows
        i0, i1 = eng.in_rows
        raw = ct.RawRaster(s2_file[:, i0:i1].contiguous().pin_memory(), s1[:, i0:i1].contiguous().pin_memory())
        host = torch.full((hi - lo, W), float("nan")).pin_memory()
        with torch.no_grad():
            o = eng.run(raw, ids[lo:hi].contiguous(), 31, row_offset=i0, map_out=host)
            eng.wait_download()
        torch.cuda.synchronize()
        outs.append((o["map"].clone(), o["count"].clone(), o["sums"].clone(), host.clone(), eng.h2d_bytes))
    a, b = outs
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[3], b[3])
    assert torch.equal(b[3], b[0].cpu())
    assert torch.allclose(a[2], b[2], rtol=1e-6)
    assert b[4] == (i1 - i0) * W * 16 and b[4] < a[4]


def test_prefetched_runs_pipeline_and_match(model):
    """prefetch() + run() + run(): consecutive rasters through double-buffered slabs, the previous map still downloading while the
    next raster computes — same maps / sums / host copies as isolated runs."""
    H, W, ps, ov = 900, 456, 128, 32
    ids = po.synthetic_regions(H, W, 30).cuda()
    rasters = []
    for seed in (35, 36, 37):
        s2_file, s1 = po.synthetic_raw(H, W, seed=seed)
        rasters.append(ct.RawRaster(s2_file.pin_memory(), s1.pin_memory()))
    ref = []
    with torch.no_grad():
        for r in rasters:
            e = ct.CountryEngine([model], H, W, ps, ov, merge=True, rows_per_strip=2, upload_once=True)
            lo, hi = e.out_rows
            o = e.run(r, ids[lo:hi].contiguous(), 31)
            ref.append((o["map"].clone(), o["sums"].clone()))
        eng = ct.CountryEngine([model], H, W, ps, ov, merge=True, rows_per_strip=3, first_strip_rows=1, last_strip_rows=1, upload_once=True)
        lo, hi = eng.out_rows
        hosts = [torch.full((hi - lo, W), float("nan")).pin_memory() for _ in rasters]
        eng.prefetch(rasters[0])
        outs = []
        for k, r in enumerate(rasters):
            o = eng.run(r, ids[lo:hi].contiguous(), 31, map_out=hosts[k])
            if k + 1 < len(rasters):
                eng.prefetch(rasters[k + 1])
            outs.append((o["map"], o["sums"]))          # no synchronisation between the rasters
        eng.wait_download()
        torch.cuda.synchronize()
    for k in range(len(rasters)):
        assert torch.equal(outs[k][0], ref[k][0]) and torch.equal(hosts[k], ref[k][0].cpu())
        assert torch.allclose(outs[k][1], ref[k][1], rtol=1e-6)
