"""GPU: tiled country inference + census aggregation against the golden map produced by the reference forward
inside the restated run_eval loop (oracle/make_golden.py), in reference-tile and merged-strip mode."""
import pytest
import torch

from popcorn_b200 import country as ct
from popcorn_b200 import ops
from oracle import popcorn_oracle as po
from util import TOL_PIXEL, TOL_REGION, build_model, golden, golden_state_dict, max_rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    return build_model(golden_state_dict()).eval()


@pytest.mark.parametrize("merge,rps", [(False, 1), (True, 1), (True, 2)])
@pytest.mark.parametrize("streamed", [False, True])
def test_tiled_eval_vs_reference_golden(model, merge, rps, streamed):
    g = golden("tiled_eval")
    ps, ov = int(g["patchsize"]), int(g["overlap"])
    raster = g["raster"]
    _, H, W = raster.shape
    R = int(g["ids"].max()) + 1
    eng = ct.CountryEngine([model], H, W, ps, ov, merge=merge, rows_per_strip=rps)
    src = raster.pin_memory() if streamed else raster.cuda()
    with torch.no_grad():
        out = eng.run(src, g["ids"].cuda().contiguous(), R)
    assert torch.equal(out["count"].cpu(), g["count"])
    assert max_rel(out["map"], g["map"]) < TOL_PIXEL
    assert max_rel(out["scale_map"], g["scale_map"]) < TOL_PIXEL
    valid = g["census"] > -1
    assert max_rel(out["sums"][1:].float().cpu()[valid], g["census"][valid], floor_frac=1.0) < TOL_REGION


def test_merged_strips_are_bit_identical_to_reference_tiles(model):
    """Merging main-grid tiles keeps every pool phase, so the written pixels do not change at all."""
    H, W, ps, ov = 520, 456, 128, 32
    raster = po.synthetic_input(H, W, seed=21)[0].cuda()
    ids = po.synthetic_regions(H, W, 30).cuda()
    outs = []
    for merge, rps in ((False, 1), (True, 1), (True, 3)):
        eng = ct.CountryEngine([model], H, W, ps, ov, merge=merge, rows_per_strip=rps)
        with torch.no_grad():
            outs.append(eng.run(raster, ids, 31))
    for o in outs[1:]:
        assert torch.equal(o["map"], outs[0]["map"])
        assert torch.equal(o["count"], outs[0]["count"])
        assert torch.allclose(o["sums"], outs[0]["sums"], rtol=1e-9)


def test_rank_sharded_partials_sum_to_the_single_gpu_result(model):
    """world_size 1 process emulating ranks 0..2: per-rank partial sums add up to the unsharded census sums."""
    H, W, ps, ov = 900, 300, 128, 32
    raster = po.synthetic_input(H, W, seed=22)[0].cuda()
    ids = po.synthetic_regions(H, W, 20).cuda()
    with torch.no_grad():
        full = ct.CountryEngine([model], H, W, ps, ov, merge=True, rows_per_strip=1).run(raster, ids, 21)
        total = torch.zeros_like(full["sums"])
        for r in range(3):
            eng = ct.CountryEngine([model], H, W, ps, ov, merge=True, rows_per_strip=1, rank=r, world=3)
            lo, hi = eng.out_rows
            i0, i1 = eng.in_rows
            o = eng.run(raster[:, i0:i1].contiguous(), ids[lo:hi].contiguous(), 21, row_offset=i0)
            assert torch.equal(o["map"], full["map"][lo:hi])
            total += o["sums"]
    assert max_rel(total, full["sums"], floor_frac=1.0) < 1e-9


def test_ensemble_mean_std_and_dasymetric_adjust(model):
    sd2 = po.random_state_dict(seed=5)
    gsd = golden_state_dict()
    for k in sd2:                       # second member: different unetmodel / head, same building_extractor
        if k.startswith("building_extractor."):
            sd2[k] = gsd[k]
    m2 = build_model(sd2).eval()
    H, W, ps, ov = 260, 300, 128, 32
    raster = po.synthetic_input(H, W, seed=23)[0]
    ids = po.synthetic_regions(H, W, 6)
    eng = ct.CountryEngine([model, m2], H, W, ps, ov, merge=True)
    assert eng._bext_shared
    with torch.no_grad():
        out = eng.run(raster.cuda(), ids.cuda(), 7)
        ref_map, ref_std, ref_scale, ref_cnt = po.tiled_eval([gsd, sd2], raster, ps, ov)
    assert torch.equal(out["count"].cpu(), ref_cnt)
    assert max_rel(out["map"], ref_map) < TOL_PIXEL
    covered = ref_cnt > 0
    assert max_rel(out["std"].cpu()[covered], ref_std[covered], floor_frac=1e-2) < 5e-2
    # dasymetric adjustment: afterwards every region with a non-zero prediction sums to its census count
    pop = torch.arange(7, dtype=torch.float32) * 1000 + 500
    adj = ct.adjust_map_to_census(out["map"].clone(), ids.cuda(), out["sums"], pop)
    sums2 = ops.region_sum(adj, ids.cuda(), 7).cpu()
    nz = out["sums"].cpu() > 0
    assert torch.allclose(sums2[nz].float(), pop[nz], rtol=1e-4)
    ref_adj = po.adjust_map_to_census(ref_map, ids.float(), list(range(7)), po.region_bboxes(ids, 7)[:0] or
                                      [(0, H, 0, W)] * 7, pop)
    assert max_rel(adj, ref_adj) < TOL_PIXEL
