"""CPU: the country / time-series engines' host logic (row offsets of sharded rasters, upload ring, upload-once slabs, strip
shipping to map_out, balanced shards, empty ranks) on a stand-in device (tests/fake_device.py: torch-CPU `ops` with a 3x3-local
fake network, no-op streams).  Expected values come from applying the same fake network to the WHOLE raster at once and the
reference tile grid's visit counts (oracle.get_patch_indices / centre_mask) — i.e. what the tiling must be invisible to."""
import pytest
import torch

import fake_device as fd
from popcorn_b200 import country as ct
from popcorn_b200 import timeseries as ts
from oracle import popcorn_oracle as po


class Pinned(torch.Tensor):
    def is_pinned(self):
        return True


class OnDevice(torch.Tensor):
    is_cuda = property(lambda self: True)


def _expected(norm, ids, H, W, ps, ov, R):
    dens, scale = fd.fake_density(norm[None])
    cnt = torch.zeros(H, W, dtype=torch.int16)
    m = po.centre_mask(ps, ps, ov)
    for xl, yl in po.get_patch_indices(H, W, ps, ov).tolist():
        cnt[xl:xl + ps, yl:yl + ps][m] += 1
    dmap = torch.where(cnt > 0, dens[0], torch.zeros(()))
    sums = torch.zeros(R, dtype=torch.float64).index_add_(0, ids.reshape(-1).long(), dmap.reshape(-1).double())
    return dmap, torch.where(cnt > 0, scale[0], torch.zeros(())), cnt, sums


@pytest.fixture()
def world_data():
    H, W, ps, ov, R = 1220, 300, 192, 32, 21
    s2_file, s1 = po.synthetic_raw(H, W, seed=41)
    norm = fd.normalise(s2_file, s1)
    ids = po.synthetic_regions(H, W, R - 1)
    return H, W, ps, ov, R, s2_file, s1, norm, ids, _expected(norm, ids, H, W, ps, ov, R)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("mode", ["device", "host_fp32", "raw_windows", "raw_once", "raw_device"])
@pytest.mark.parametrize("balance", [False, True])
def test_sharded_engine_reassembles_the_whole_raster_result(monkeypatch, world_data, world, mode, balance):
    H, W, ps, ov, R, s2_file, s1, norm, ids, (want_map, want_scale, want_cnt, want_sums) = world_data
    if balance and world == 1:
        pytest.skip("nothing to balance")
    log = fd.install(monkeypatch)
    got_map = torch.zeros(H, W)
    got_cnt = torch.zeros(H, W, dtype=torch.int16)
    host_full = torch.zeros(H, W)
    total = torch.zeros(R, dtype=torch.float64)
    seen = []
    for rank in range(world):
        eng = ct.CountryEngine([fd.FakeModel()], H, W, ps, ov, merge=True, rows_per_strip=2, rank=rank, world=world,
                               first_strip_rows=1, balance=balance, balance_unit=64, upload_once=(mode == "raw_once"))
        lo, hi = eng.out_rows
        i0, i1 = eng.in_rows
        if mode == "device":
            raster = norm[:, i0:i1].contiguous().as_subclass(OnDevice)
        elif mode == "host_fp32":
            raster = norm[:, i0:i1].contiguous().as_subclass(Pinned)
        else:
            raster = ct.RawRaster(s2_file[:, i0:i1].contiguous(), s1[:, i0:i1].contiguous())
            if mode == "raw_device":
                raster.is_cuda = True
        map_out = torch.full((hi - lo, W), float("nan")).as_subclass(Pinned) if mode != "device" else None
        del log[:]
        out = eng.run(raster, ids[lo:hi].contiguous(), R, row_offset=i0, map_out=map_out)
        eng.wait_download()
        if not eng.windows:
            assert (lo, hi) == (0, 0) and float(out["sums"].abs().sum()) == 0.0 and out["map"].numel() == 0
            continue
        seen.append((lo, hi))
        got_map[lo:hi] = out["map"]
        got_cnt[lo:hi] = out["count"]
        total += out["sums"]
        if map_out is not None:
            assert torch.equal(torch.Tensor(map_out), out["map"]), "rows shipped to the host differ from the device map"
            host_full[lo:hi] = map_out
            shipped = sorted((e[1], e[2]) for e in log if e[0] == "finalize")
            assert shipped[0][0] == 0 and shipped[-1][1] == hi - lo
            assert all(a[1] == b[0] for a, b in zip(shipped, shipped[1:])), "every owned row is finalised exactly once"
        h2d = sum(e[1] for e in log if e[0] == "h2d")
        if mode == "raw_once":
            assert h2d == eng.h2d_bytes == (i1 - i0) * W * 16
        elif mode == "raw_windows":
            assert h2d == eng.h2d_bytes == sum(w.h * w.w for w in eng.windows) * 16 >= (i1 - i0) * W * 16
        elif mode == "host_fp32":
            assert h2d == eng.h2d_bytes == sum(w.h * w.w for w in eng.windows) * 24
    assert all(a[1] <= b[0] for a, b in zip(seen, seen[1:]))
    assert torch.equal(got_cnt, want_cnt)
    assert torch.allclose(got_map, want_map, rtol=1e-6, atol=1e-6)
    assert torch.allclose(total, want_sums, rtol=1e-9)
    if mode != "device":
        assert torch.equal(host_full, got_map)


def test_time_series_engine_on_sharded_frames(monkeypatch, world_data):
    H, W, ps, ov, R, s2_file, s1, norm, ids, (want_map, _, _, want_sums) = world_data
    fd.install(monkeypatch)
    frames_full = [norm, norm * 0.5, norm + 0.25]
    wants = [_expected(f, ids, H, W, ps, ov, R) for f in frames_full]
    season = sum(w[0] for w in wants) / 3
    world = 4
    tot = torch.zeros(4, dtype=torch.float64)
    for rank in range(world):
        eng = ts.TimeSeriesEngine([fd.FakeModel()], H, W, rank=rank, world=world, patch=ps, overlap=ov, merge=True, rows_per_strip=2)
        lo, hi = eng.out_rows
        i0, i1 = eng.in_rows
        frames = [f[:, i0:i1].contiguous().as_subclass(OnDevice) for f in frames_full]
        o = eng.run(frames, ids[lo:hi].contiguous(), R, row_offset=i0)
        assert torch.allclose(o["season_map"], season[lo:hi], rtol=1e-5, atol=1e-6)
        tot[:3] += o["totals"]
        tot[3] += o["season_total"]
    for t in range(3):
        assert abs(float(tot[t]) - float(wants[t][0].double().sum())) < 1e-6 * float(wants[t][0].double().sum())
    assert abs(float(tot[3]) - float(season.double().sum())) < 1e-6 * float(season.double().sum())


def test_plan_frames_uses_every_rank():
    for T, world in [(4, 1), (4, 2), (4, 4), (4, 8), (3, 8), (4, 6), (1, 8), (5, 2)]:
        F, S, plan = ts.plan_frames(T, world)
        assert F * S == world and F <= T and len(plan) == world
        assert all(p[2] for p in plan), "every rank owns at least one frame"
        for t in range(T):     # every frame is covered by exactly the S row shards of one frame group
            owners = [(g, k) for g, k, fr in plan if t in fr]
            assert sorted(k for _, k in owners) == list(range(S)) and len({g for g, _ in owners}) == 1
    assert ts.plan_frames(4, 8)[:2] == (4, 2)      # BASELINE config 5: 4 seasonal frames on 8 GPUs -> 4 frame groups x 2 row shards


def _ts_frames_worker(rank, world, port, T, q):
    """One rank of a frames x row-strips time series on the stand-in device, real torch.distributed (gloo) collectives."""
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mp_ = pytest.MonkeyPatch()
    try:
        fd.install(mp_)
        H, W, ps, ov, R = 1220, 300, 192, 32, 21
        s2_file, s1 = po.synthetic_raw(H, W, seed=41)
        norm = fd.normalise(s2_file, s1)
        ids = po.synthetic_regions(H, W, R - 1)
        frames_full = [norm * (1.0 - 0.1 * t) + 0.05 * t for t in range(T)]
        eng = ts.TimeSeriesEngine([fd.FakeModel()], H, W, rank=rank, world=world, frames=T, patch=ps, overlap=ov, merge=True,
                                  rows_per_strip=2)
        lo, hi = eng.out_rows
        i0, i1 = eng.in_rows
        frames = {t: frames_full[t][:, i0:i1].contiguous().as_subclass(OnDevice) for t in eng.my_frames}
        o = eng.run(frames, ids[lo:hi].contiguous(), R, row_offset=i0)
        q.put((rank, eng.F, eng.S, eng.my_frames, (lo, hi), torch.Tensor(o["season_map"]).clone(), o["totals"].clone(),
               float(o["season_total"]), o["sums"].clone()))
    finally:
        mp_.undo()
        dist.destroy_process_group()


@pytest.mark.parametrize("world,T", [(2, 4), (4, 2), (4, 4)])
def test_time_series_frames_x_row_strips_over_gloo(world, T):
    """BASELINE config 5's partition (SURVEY.md §8e second axis): frame groups x row shards with the real collectives — census sums
    inside a frame group, season map across the groups, totals over everyone — against the unsharded result."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ts_frames_worker, args=(r, world, port, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    H, W, ps, ov, R = 1220, 300, 192, 32, 21
    s2_file, s1 = po.synthetic_raw(H, W, seed=41)
    norm = fd.normalise(s2_file, s1)
    ids = po.synthetic_regions(H, W, R - 1)
    wants = [_expected(norm * (1.0 - 0.1 * t) + 0.05 * t, ids, H, W, ps, ov, R) for t in range(T)]
    season = sum(w[0] for w in wants) / T
    F, S, plan = ts.plan_frames(T, world)
    for rank, f, s_, mine, (lo, hi), smap, totals, stot, sums in res:
        assert (f, s_) == (F, S) and mine == plan[rank][2]
        assert torch.allclose(smap, season[lo:hi], rtol=1e-5, atol=1e-6)          # every rank ends with the full season average of its rows
        for t in range(T):
            want = float(wants[t][0].double().sum())
            assert abs(float(totals[t]) - want) < 1e-6 * want
            assert torch.allclose(sums[t], wants[t][3], rtol=1e-9)
        assert abs(stot - float(season.double().sum())) < 1e-6 * float(season.double().sum())


def test_prefetch_pipelines_consecutive_rasters(monkeypatch, world_data):
    """CountryEngine.prefetch(): the next raster is uploaded into the other slab set while the current one is (logically) still
    computing; run() on the prefetched object uploads nothing again and gives the result of a fresh engine."""
    H, W, ps, ov, R, s2_file, s1, norm, ids, (want_map, _, want_cnt, want_sums) = world_data
    log = fd.install(monkeypatch)
    eng = ct.CountryEngine([fd.FakeModel()], H, W, ps, ov, merge=True, rows_per_strip=3, first_strip_rows=1, last_strip_rows=1,
                           upload_once=True)
    lo, hi = eng.out_rows
    a = ct.RawRaster(s2_file.clone(), s1.clone())
    b = ct.RawRaster(s2_file.clone(), (s1 * 0.5).contiguous())
    out_a = eng.run(a, ids[lo:hi].contiguous(), R)
    map_a, sums_a = out_a["map"].clone(), out_a["sums"].clone()
    assert torch.allclose(map_a, want_map[lo:hi], rtol=1e-6, atol=1e-6) and torch.allclose(sums_a, want_sums, rtol=1e-9)
    del log[:]
    eng.prefetch(b)
    up = sum(e[1] for e in log if e[0] == "h2d")
    assert up == H * W * 16                       # the whole raster went up at prefetch time, into the OTHER slab set
    del log[:]
    out_b = eng.run(b, ids[lo:hi].contiguous(), R)
    assert sum(e[1] for e in log if e[0] == "h2d") == 0, "run() must use the prefetched slabs"
    fresh = ct.CountryEngine([fd.FakeModel()], H, W, ps, ov, merge=True, rows_per_strip=2, upload_once=True)
    ref_b = fresh.run(ct.RawRaster(s2_file.clone(), (s1 * 0.5).contiguous()), ids[lo:hi].contiguous(), R)
    assert torch.equal(out_b["map"], ref_b["map"]) and torch.equal(out_b["count"], ref_b["count"])
    assert torch.allclose(out_b["sums"], ref_b["sums"], rtol=1e-12)
    # a raster that was not prefetched is uploaded by run() itself, again into the set that is free
    del log[:]
    out_a2 = eng.run(a, ids[lo:hi].contiguous(), R)
    assert sum(e[1] for e in log if e[0] == "h2d") == H * W * 16
    assert torch.equal(out_a2["map"], map_a)
