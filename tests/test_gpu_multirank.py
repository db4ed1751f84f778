"""GPU, >= 2 devices: the N > 1 paths with REAL NCCL ranks (one process per GPU) against the single-GPU result —
row-sharded country inference (BASELINE config 4: balanced shards, one all-reduce of the census sums) and the frames x row-strips
time series (config 5: census sums inside a frame group, season map across the groups, totals over everyone).
Skipped on a one-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py` runs it."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from popcorn_b200 import country as ct
        from popcorn_b200 import timeseries as ts
        from oracle import popcorn_oracle as po
        from util import build_model, golden_state_dict
        model = build_model(golden_state_dict(), device=f"cuda:{rank}").eval()
        H, W, ps, ov, R = 1220, 300, 192, 32, 21
        s2, s1 = po.synthetic_raw(H, W, seed=35)
        ids = po.synthetic_regions(H, W, R - 1).cuda()
        res = {}
        with torch.no_grad():
            # ---- config 4: rows sharded (balanced 64-row units), raw host input uploaded once, map shipped to the host
            eng = ct.CountryEngine([model], H, W, ps, ov, merge=True, rows_per_strip=2, rank=rank, world=world, balance=True,
                                   balance_unit=64, upload_once=True)
            lo, hi = eng.out_rows
            i0, i1 = eng.in_rows
            raw = ct.RawRaster(s2[:, i0:i1].contiguous().pin_memory(), s1[:, i0:i1].contiguous().pin_memory())
            host = torch.zeros(hi - lo, W).pin_memory()
            out = eng.run(raw, ids[lo:hi].contiguous(), R, row_offset=i0, map_out=host)
            eng.wait_download()
            torch.cuda.synchronize()
            res["country"] = ((lo, hi), host.clone(), out["sums"].cpu(), out["count"].cpu())
            # ---- config 5: 4 seasonal frames, frames x row strips
            T = 4
            frames_full = [po.read_and_normalize(*po.synthetic_raw(H, W, seed=70 + t))[0] for t in range(T)]
            tse = ts.TimeSeriesEngine([model], H, W, rank=rank, world=world, frames=T, patch=ps, overlap=ov, merge=True, rows_per_strip=2)
            lo2, hi2 = tse.out_rows
            j0, j1 = tse.in_rows
            frames = {t: frames_full[t][:, j0:j1].contiguous().cuda() for t in tse.my_frames}
            o = tse.run(frames, ids[lo2:hi2].contiguous(), R, row_offset=j0)
            res["series"] = ((lo2, hi2), tse.describe(), tse.my_frames, o["season_map"].cpu(), o["totals"].cpu(),
                             float(o["season_total"]), o["sums"].cpu())
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (real NCCL ranks)")
def test_two_nccl_ranks_match_the_single_gpu_result():
    import torch.multiprocessing as mp
    from popcorn_b200 import country as ct
    from popcorn_b200 import timeseries as ts
    from oracle import popcorn_oracle as po
    from util import build_model, golden_state_dict, max_rel
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single-GPU references
    model = build_model(golden_state_dict()).eval()
    H, W, ps, ov, R = 1220, 300, 192, 32, 21
    s2, s1 = po.synthetic_raw(H, W, seed=35)
    ids = po.synthetic_regions(H, W, R - 1).cuda()
    with torch.no_grad():
        e0 = ct.CountryEngine([model], H, W, ps, ov, merge=True, rows_per_strip=2)
        flo, fhi = e0.out_rows
        full = e0.run(ct.RawRaster(s2.cuda(), s1.cuda()), ids[flo:fhi].contiguous(), R)
        T = 4
        frames_full = [po.read_and_normalize(*po.synthetic_raw(H, W, seed=70 + t))[0].cuda() for t in range(T)]
        t0 = ts.TimeSeriesEngine([model], H, W, patch=ps, overlap=ov, merge=True, rows_per_strip=2)
        sfull = t0.run(frames_full, ids[flo:fhi].contiguous(), R)
    seen = 0
    for rank in range(world):
        (lo, hi), host, sums, cnt = got[rank]["country"]
        assert torch.equal(host, full["map"][lo - flo:hi - flo].cpu())          # bit-identical rows, shipped through map_out
        assert torch.equal(cnt, full["count"][lo - flo:hi - flo].cpu())
        # all-reduced: every rank holds the country's sums (fp32 per-CTA bins before the fp64 atomics: order-dependent in the last bits)
        assert max_rel(sums, full["sums"], floor_frac=1.0) < 1e-6
        seen += hi - lo
        (lo2, hi2), desc, mine, smap, totals, stot, ssums = got[rank]["series"]
        assert desc == "2 frame group(s) x 1 row shard(s)" and mine == [rank, rank + 2]
        assert max_rel(smap, sfull["season_map"][lo2 - flo:hi2 - flo], floor_frac=1.0) < 1e-6
        assert max_rel(totals, sfull["totals"], floor_frac=1.0) < 1e-6
        assert abs(stot - float(sfull["season_total"])) < 1e-6 * float(sfull["season_total"])
        assert max_rel(ssums, sfull["sums"], floor_frac=1.0) < 1e-6
    assert seen == fhi - flo
