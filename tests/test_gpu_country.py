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
    assert eng.merged == merge
    lo, hi = eng.out_rows                         # the outer `overlap` frame of the raster is never written (stays 0)
    assert (lo, hi) == (ov, H - ov) and float(g["map"][:lo].abs().sum() + g["map"][hi:].abs().sum()) == 0.0
    src = raster.pin_memory() if streamed else raster.cuda()
    with torch.no_grad():
        out = eng.run(src, g["ids"][lo:hi].cuda().contiguous(), R)
    assert torch.equal(out["count"].cpu(), g["count"][lo:hi])
    assert max_rel(out["map"], g["map"][lo:hi]) < TOL_PIXEL
    assert max_rel(out["scale_map"], g["scale_map"][lo:hi]) < TOL_PIXEL
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
        lo, hi = eng.out_rows
        with torch.no_grad():
            outs.append(eng.run(raster, ids[lo:hi].contiguous(), 31))
    for o in outs[1:]:
        assert torch.equal(o["map"], outs[0]["map"])
        assert torch.equal(o["count"], outs[0]["count"])
        assert torch.allclose(o["sums"], outs[0]["sums"], rtol=1e-6)   # per-CTA fp32 bins: atomic order varies


def test_rank_sharded_partials_sum_to_the_single_gpu_result(model):
    """world_size 1 process emulating ranks 0..2: per-rank partial sums add up to the unsharded census sums."""
    H, W, ps, ov = 900, 300, 128, 32
    raster = po.synthetic_input(H, W, seed=22)[0].cuda()
    ids = po.synthetic_regions(H, W, 20).cuda()
    with torch.no_grad():
        e0 = ct.CountryEngine([model], H, W, ps, ov, merge=True, rows_per_strip=1)
        flo, fhi = e0.out_rows
        full = e0.run(raster, ids[flo:fhi].contiguous(), 21)
        total = torch.zeros_like(full["sums"])
        for r in range(3):
            eng = ct.CountryEngine([model], H, W, ps, ov, merge=True, rows_per_strip=1, rank=r, world=3)
            lo, hi = eng.out_rows
            i0, i1 = eng.in_rows
            o = eng.run(raster[:, i0:i1].contiguous(), ids[lo:hi].contiguous(), 21, row_offset=i0)
            assert torch.equal(o["map"], full["map"][lo - flo:hi - flo])
            total += o["sums"]
    assert max_rel(total, full["sums"], floor_frac=1.0) < 1e-6   # fp32 warp partials regroup across shards


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
    lo, hi = eng.out_rows
    ids_own = ids[lo:hi].cuda().contiguous()
    with torch.no_grad():
        out = eng.run(raster.cuda(), ids_own, 7)
        ref_map, ref_std, ref_scale, ref_cnt = po.tiled_eval([gsd, sd2], raster, ps, ov)
    assert torch.equal(out["count"].cpu(), ref_cnt[lo:hi])
    assert max_rel(out["map"], ref_map[lo:hi]) < TOL_PIXEL
    covered = ref_cnt[lo:hi] > 0
    assert max_rel(out["std"].cpu()[covered], ref_std[lo:hi][covered], floor_frac=1e-2) < 5e-2
    # dasymetric adjustment: afterwards every region with a non-zero prediction sums to its census count
    pop = torch.arange(7, dtype=torch.float32) * 1000 + 500
    adj = ct.adjust_map_to_census(out["map"].clone(), ids_own, out["sums"], pop)
    sums2 = ops.region_sum(adj, ids_own, 7).cpu()
    nz = out["sums"].cpu() > 0
    assert torch.allclose(sums2[nz].float(), pop[nz], rtol=1e-4)
    ref_adj = po.adjust_map_to_census(ref_map, ids.float(), list(range(7)), [(0, H, 0, W)] * 7, pop)
    assert max_rel(adj, ref_adj[lo:hi]) < TOL_PIXEL


def test_merging_is_refused_below_the_receptive_field(model):
    """overlap 16 < 24: tile-border padding artefacts reach the written centre, so windows must not be merged."""
    eng = ct.CountryEngine([model], 200, 236, 96, 16, merge=True)
    assert not eng.merged and all(w.h == 96 and w.w == 96 for w in eng.windows)


@pytest.mark.parametrize("shape", [(64, 64), (37, 101), (130, 259)])
def test_raw_ingest_is_bit_identical_to_reference_normalisation(shape):
    """pc_ingest_normalize (uint16 S2 in file band order + float32 S1) == the reference's .astype(float32) + apply_normalize
    + concatenate, restated in oracle.read_and_normalize (pinned to utils/utils.py in tests/test_oracle_vs_reference.py)."""
    s2_file, s1 = po.synthetic_raw(*shape, seed=shape[0])
    want = po.read_and_normalize(s2_file, s1)[0]
    got = ops.ingest_normalize(s2_file.cuda(), s1.cuda(), s2_plane_map=ops.S2_FILE_TO_RGBN)
    assert torch.equal(got.cpu(), want)
    # a strided window of a bigger raster, float32 S2 variant, identity band order
    H, W = shape
    if H > 40 and W > 40:
        big2 = s2_file[[2, 1, 0, 3]].float().cuda()
        got = ops.ingest_normalize(big2[:, 3:H - 5, 7:W - 2], s1.cuda()[:, 3:H - 5, 7:W - 2], s2_plane_map=ops.S2_IDENTITY)
        assert torch.equal(got.cpu(), want[:, 3:H - 5, 7:W - 2])


@pytest.mark.parametrize("streamed", [False, True])
def test_country_engine_on_raw_rasters_equals_normalised_input(model, streamed):
    """RawRaster path (16 B/px upload + device-side normalisation) gives exactly the map of the fp32 path fed with the
    reference-normalised raster."""
    H, W, ps, ov = 520, 456, 128, 32
    s2_file, s1 = po.synthetic_raw(H, W, seed=33)
    norm = po.read_and_normalize(s2_file, s1)[0]
    ids = po.synthetic_regions(H, W, 30).cuda()
    eng = ct.CountryEngine([model], H, W, ps, ov, merge=True, rows_per_strip=2)
    lo, hi = eng.out_rows
    with torch.no_grad():
        a = eng.run(norm.cuda(), ids[lo:hi].contiguous(), 31)
        a = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in a.items()}
        raw = ct.RawRaster(s2_file.pin_memory(), s1.pin_memory()) if streamed else ct.RawRaster(s2_file.cuda(), s1.cuda())
        b = eng.run(raw, ids[lo:hi].contiguous(), 31)
    assert torch.equal(a["map"], b["map"]) and torch.equal(a["count"], b["count"])
    assert torch.allclose(a["sums"], b["sums"], rtol=1e-6)
    if streamed:
        assert eng.h2d_bytes == sum(w.h * w.w for w in eng.windows) * 16


def test_map_out_streams_the_finalised_map_to_the_host(model):
    """run(map_out=pinned) ships finished strips while later strips compute; the host copy equals the device map."""
    H, W, ps, ov = 700, 456, 128, 32
    raster = po.synthetic_input(H, W, seed=5)[0]
    ids = po.synthetic_regions(H, W, 30).cuda()
    eng = ct.CountryEngine([model], H, W, ps, ov, merge=True, rows_per_strip=2)
    lo, hi = eng.out_rows
    with torch.no_grad():
        ref = eng.run(raster.cuda(), ids[lo:hi].contiguous(), 31)
        ref_map, ref_sums = ref["map"].cpu(), ref["sums"].cpu()
        host = torch.full((hi - lo, W), float("nan")).pin_memory()
        out = eng.run(raster.pin_memory(), ids[lo:hi].contiguous(), 31, map_out=host)
        eng.wait_download()
    assert torch.equal(host, ref_map) and torch.equal(out["map"].cpu(), ref_map)
    assert torch.allclose(out["sums"].cpu(), ref_sums, rtol=1e-6)
