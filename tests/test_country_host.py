"""CPU: window planning / sharding logic of the country driver and the N>1 combine step (gloo, world_size 2)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from popcorn_b200 import country as ct
from oracle import popcorn_oracle as po


def _count_from_windows(wins, H, W, ov):
    c = torch.zeros(H, W, dtype=torch.int16)
    for w in wins:
        c[w.y0 + ov: w.y0 + w.h - ov, w.x0 + ov: w.x0 + w.w - ov] += 1
    return c


def _count_reference(H, W, ps, ov):
    c = torch.zeros(H, W, dtype=torch.int16)
    m = po.centre_mask(ps, ps, ov)
    for xl, yl in po.get_patch_indices(H, W, ps, ov).tolist():
        c[xl:xl + ps, yl:yl + ps][m] += 1
    return c


@pytest.mark.parametrize("H,W,ps,ov", [(200, 236, 96, 16), (300, 97, 96, 16), (97, 400, 96, 16), (1000, 1100, 256, 32),
                                       (15104 // 8, 17216 // 8, 256, 16)])
@pytest.mark.parametrize("merge,rps", [(False, 1), (True, 1), (True, 2), (True, 3)])
def test_windows_cover_exactly_like_the_reference_tile_grid(H, W, ps, ov, merge, rps):
    wins = ct.plan_windows(H, W, ps, ov, merge, rps)
    assert torch.equal(_count_from_windows(wins, H, W, ov), _count_reference(H, W, ps, ov))
    assert sum(w.ntiles for w in wins) == len(po.get_patch_indices(H, W, ps, ov))
    stride = ps - 2 * ov
    for w in wins:   # merged windows keep every tile's pool phase: origins on the reference grid, size = k*stride + 2*ov
        assert (w.h - 2 * ov) % stride == 0 and (w.w - 2 * ov) % stride == 0
        assert w.y0 % stride == 0 or w.y0 == H - ps
        assert w.x0 % stride == 0 or w.x0 == W - ps


def test_unmerged_plan_is_the_reference_index_list():
    H, W, ps, ov = 200, 236, 96, 16
    mine = sorted((w.y0, w.x0) for w in ct.plan_windows(H, W, ps, ov, merge=False))
    assert mine == sorted(map(tuple, po.get_patch_indices(H, W, ps, ov).tolist()))


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("rps", [1, 2])
def test_sharding_partitions_windows_and_keeps_contributors_together(world, rps):
    H, W, ps, ov = 1900, 700, 256, 32
    wins = ct.plan_windows(H, W, ps, ov, True, rps)
    n_rows = len(ct.grid_origins(H, ps, ov))
    parts = [ct.shard_windows(wins, n_rows, r, world, rps) for r in range(world)]
    assert sorted(sum(parts, []), key=lambda w: (w.y0, w.x0)) == sorted(wins, key=lambda w: (w.y0, w.x0))
    owner = torch.full((H, W), -1, dtype=torch.int16)
    for r, part in enumerate(parts):
        for w in part:
            reg = owner[w.y0 + ov: w.y0 + w.h - ov, w.x0 + ov: w.x0 + w.w - ov]
            assert bool(((reg == -1) | (reg == r)).all()), "a pixel would be accumulated on two ranks"
            reg.fill_(r)
        lo, hi = ct.owned_rows(part, H, ov)
        if part:
            assert all(lo <= w.y0 + ov and w.y0 + w.h - ov <= hi for w in part)
    # row ownership ranges of different ranks do not overlap
    ranges = sorted(ct.owned_rows(p, H, ov) for p in parts if p)
    assert all(a[1] <= b[0] for a, b in zip(ranges, ranges[1:]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, R, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ps, ov = 96, 16
        g = torch.Generator().manual_seed(5)
        full_map = torch.rand(H, W, generator=g)
        ids = po.synthetic_regions(H, W, R, seed=2)
        wins = ct.plan_windows(H, W, ps, ov, True, 1)
        mine = ct.shard_windows(wins, len(ct.grid_origins(H, ps, ov)), rank, world, 1)
        lo, hi = ct.owned_rows(mine, H, ov)
        # each rank "computes" only the pixels its windows write; partial census sums over its own rows
        local = torch.zeros(H, W)
        for w in mine:
            ys, xs = slice(w.y0 + ov, w.y0 + w.h - ov), slice(w.x0 + ov, w.x0 + w.w - ov)
            local[ys, xs] = full_map[ys, xs]
        part = po.region_sums(local[lo:hi], ids[lo:hi], R + 1)
        total = ct.allreduce_sums(part.clone())
        q.put((rank, total))
    finally:
        dist.destroy_process_group()


def test_two_rank_partial_sums_allreduce_gloo():
    H, W, R = 500, 236, 9
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, H, W, R, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(5)
    full_map = torch.rand(H, W, generator=g)
    ids = po.synthetic_regions(H, W, R, seed=2)
    covered = _count_reference(H, W, 96, 16) > 0
    ref = po.region_sums(full_map * covered, ids, R + 1)
    assert torch.allclose(res[0], ref, rtol=1e-12) and torch.equal(res[0], res[1])


def _ts_worker(rank, world, port, q):
    """Season totals of row-sharded frames: each rank holds its rows of T frame maps; totals are all-reduced (gloo)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        T, H, W = 4, 64, 40
        g = torch.Generator().manual_seed(11)
        frames = torch.rand(T, H, W, generator=g, dtype=torch.float64)
        lo, hi = (0, 40) if rank == 0 else (40, 64)
        totals = torch.zeros(T + 1, dtype=torch.float64)
        season = torch.zeros(hi - lo, W, dtype=torch.float64)
        for t in range(T):
            totals[t] = frames[t, lo:hi].sum()
            season += frames[t, lo:hi]
        season /= T
        totals[T] = season.sum()
        ct.allreduce_sums(totals)
        q.put((rank, totals))
    finally:
        dist.destroy_process_group()


def test_time_series_totals_allreduce_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ts_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(11)
    frames = torch.rand(4, 64, 40, generator=g, dtype=torch.float64)
    want = torch.cat([frames.sum((1, 2)), frames.mean(0).sum()[None]])
    assert torch.allclose(res[0], want, rtol=1e-12) and torch.equal(res[0], res[1])


@pytest.mark.parametrize("H,W", [(4 * 1792 + 300, 3 * 1792 + 500), (9000, 5000)])
def test_short_first_strip_covers_the_same_pixels(H, W):
    """first_strip_rows only regroups main-grid tile rows: the union of written centres is unchanged."""
    a = ct.plan_windows(H, W, merge=True, rows_per_strip=3)
    b = ct.plan_windows(H, W, merge=True, rows_per_strip=3, first_strip_rows=1)
    assert b[0].h == ct.PATCH and sum(w.ntiles for w in a) == sum(w.ntiles for w in b)

    def cover(wins):
        m = torch.zeros(H // 8 + 1, W // 8 + 1, dtype=torch.int16)
        for w in wins:
            ov = ct.OVERLAP
            m[(w.y0 + ov) // 8:(w.y0 + w.h - ov) // 8, (w.x0 + ov) // 8:(w.x0 + w.w - ov) // 8] += 1
        return m
    assert torch.equal(cover(a), cover(b))


def test_random_rasters_plan_and_shard_invariants():
    """Property test (hypothesis): for random raster sizes, tile geometry, strip heights and world sizes the merged plan
    visits every pixel exactly as often as the reference tile grid, and the rank shards partition the windows with disjoint
    owned row ranges."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(st.integers(0, 3), st.integers(1, 9), st.integers(1, 9), st.integers(0, 500), st.integers(0, 500),
           st.integers(1, 4), st.integers(1, 5), st.booleans())
    def check(ovi, ny, nx, dy, dx, rps, world, short_first):
        ps = 96
        ov = (12, 16, 20, 24)[ovi]
        stride = ps - 2 * ov
        H, W = ps + ny * stride + dy % stride, ps + nx * stride + dx % stride
        ref = _count_reference(H, W, ps, ov)
        wins = ct.plan_windows(H, W, ps, ov, True, rps, 1 if short_first else None)
        cnt = torch.zeros(H, W, dtype=torch.int16)
        for w in wins:
            cnt[w.y0 + ov: w.y0 + w.h - ov, w.x0 + ov: w.x0 + w.w - ov] += 1
        assert torch.equal(cnt, ref)
        parts = [ct.shard_windows(wins, len(ct.grid_origins(H, ps, ov)), r, world, rps) for r in range(world)]
        assert sorted((w.y0, w.x0, w.h, w.w) for p in parts for w in p) == sorted((w.y0, w.x0, w.h, w.w) for w in wins)
        ranges = sorted(ct.owned_rows(p, H, ov) for p in parts if p)
        assert all(a[1] <= b[0] for a, b in zip(ranges, ranges[1:]))

    check()


def test_more_ranks_than_strips_leaves_empty_ranks_consistent():
    """Switzerland-shaped raster (config 5) on 8 ranks: 4 strips -> ranks 4..7 own nothing; their row ranges are empty, the
    non-empty shards still partition the plan (the engine then skips finalise / region sums on the empty ranks but joins the
    all-reduce)."""
    H, W = 13408, 30592
    wins = ct.plan_windows(H, W, merge=True, rows_per_strip=2)
    n_rows = len(ct.grid_origins(H))
    parts = [ct.shard_windows(wins, n_rows, r, 8, 2) for r in range(8)]
    assert [bool(p) for p in parts] == [True] * 4 + [False] * 4
    assert all(ct.owned_rows(p, H) == (0, 0) and ct.input_rows(p) == (0, 0) for p in parts[4:])
    assert sorted((w.y0, w.x0) for p in parts for w in p) == sorted((w.y0, w.x0) for w in wins)
    assert any(w.tile_row < 0 for w in parts[3]) and not any(w.tile_row < 0 for p in parts[:3] for w in p)


@pytest.mark.parametrize("H,W,ps,ov,unit", [(1900, 700, 256, 32, 64), (15104 // 8, 17216 // 8, 256, 32, 64), (1000, 1100, 320, 32, 128),
                                            (256, 300, 256, 32, 64), (448, 256, 256, 32, 64)])
@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("rps,short_first", [(1, False), (2, True)])
def test_balanced_shards_cover_exactly_like_the_reference_tile_grid(H, W, ps, ov, unit, world, rps, short_first):
    shards = ct.plan_balanced_shards(H, W, world, ps, ov, rps, unit, 1 if short_first else None)
    assert len(shards) == world
    cnt = torch.zeros(H, W, dtype=torch.int16)
    owner = torch.full((H,), -1, dtype=torch.int16)
    stride = ps - 2 * ov
    for r, part in enumerate(shards):
        cnt += _count_from_windows(part, H, W, ov)
        for w in part:
            rows = owner[w.y0 + ov: w.y0 + w.h - ov]
            assert bool(((rows == -1) | (rows == r)).all()), "a raster row would be accumulated on two ranks"
            rows.fill_(r)
            if w.tile_row >= 0:       # pool phase / CTA tiling of the reference grid, bounded window height
                assert w.y0 % unit == 0 and (w.h - 2 * ov) % unit == 0 and w.h - 2 * ov <= max(rps * stride, unit)
            else:
                assert w.y0 == H - ps and r == (world - 1 if H > ps else 0)
            assert w.x0 == 0 or w.x0 == W - ps
    assert torch.equal(cnt, _count_reference(H, W, ps, ov))
    ranges = [ct.owned_rows(p, H, ov) for p in shards if p]
    assert all(a[1] <= b[0] for a, b in zip(ranges, ranges[1:]))            # contiguous, in rank order


def test_balanced_shards_are_better_balanced_than_strips_on_the_bench_rasters():
    for world, H in ((8, 50048), (4, 25024), (2, 12512)):
        W = 47952
        work = lambda wins: sum(w.h * w.w for w in wins)
        bal = [work(p) for p in ct.plan_balanced_shards(H, W, world, first_strip_rows=1)]
        plan = ct.plan_windows(H, W, merge=True, rows_per_strip=2, first_strip_rows=1)
        strips = [work(ct.shard_windows(plan, 0, r, world, 2)) for r in range(world)]
        assert max(bal) < 0.95 * max(strips)
        assert max(bal) < 1.06 * sum(bal) / world


def test_balanced_shards_reject_unaligned_units():
    with pytest.raises(ValueError):
        ct.plan_balanced_shards(4000, 4000, 2, unit=100)         # 1792 % 100 != 0
    with pytest.raises(ValueError):
        ct.plan_balanced_shards(1000, 4000, 2)                    # smaller than the patch
