"""CPU, only where /root/reference exists (not on the GPU box): oracle == imported reference on fresh inputs."""
import pytest
import torch

from oracle import popcorn_oracle as po
from oracle import reference_shim as rs
from util import max_rel

pytestmark = pytest.mark.skipif(not rs.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref_model():
    return rs.build_reference_model(seed=1601).eval()


@pytest.mark.parametrize("shape,padding", [((64, 64), False), ((45, 83), False), ((40, 56), True)])
def test_dense_forward(ref_model, shape, padding):
    sd = {k: v.detach().clone() for k, v in ref_model.state_dict().items()}
    x = po.synthetic_input(*shape, seed=5)
    a, b = {"input": x.clone()}, {"input": x.clone()}
    with torch.no_grad():
        ref = rs.reference_forward(ref_model, a, padding=padding)
        ora = po.forward(sd, b, padding=padding)
    assert max_rel(b["building_counts"], a["building_counts"]) < 1e-3
    assert max_rel(ora["popdensemap"], ref["popdensemap"]) < 1e-3


def test_sparse_mask_is_region(ref_model):
    """with sigmoid scores builtup>0 is always true, so the index set equals admin==idx."""
    sd = {k: v.detach().clone() for k, v in ref_model.state_dict().items()}
    x = po.synthetic_input(64, 64, seed=9)
    admin = torch.zeros(1, 64, 64)
    admin[0, 10:40, 20:50] = 3
    inp = {"input": x, "admin_mask": admin, "census_idx": torch.tensor([3])}
    with torch.no_grad():
        out = po.forward(sd, inp, padding=False, sparse=True)
    assert torch.equal(out["mask"], admin == 3)


def test_read_and_normalize_matches_reference_apply_normalize():
    """oracle.read_and_normalize == the reference's own apply_normalize + concatenation (utils/utils.py:105-127, 162-171)
    on the band-reordered float32 cast of the raw window (PopulationDataset.py:565-567, 594-604), with the reference's
    dataset_stats.json.  (Tensor.cuda is a no-op here: apply_normalize calls .cuda() on the statistics.)"""
    import json
    import os
    rs.load_reference()
    import utils.utils as ru                      # the reference module (sys.path set by the shim)
    with open(os.path.join(rs.REF_ROOT, "data", "config", "dataset_stats.json")) as f:
        stats = json.load(f)
    for k in stats:
        for kk in stats[k]:
            stats[k][kk] = torch.tensor(stats[k][kk])
    assert tuple(float("%.4f" % v) for v in stats["sen2springNIR"]["mean"]) == po.REF_STATS["sen2springNIR"]["mean"]
    s2_file, s1 = po.synthetic_raw(40, 56, seed=3)
    sample = {"S2": s2_file[[2, 1, 0, 3]].float()[None], "S1": s1.float()[None]}
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        ref = ru.apply_transformations_and_normalize(sample, None, stats)["input"]
    finally:
        torch.Tensor.cuda = orig
    assert torch.equal(po.read_and_normalize(s2_file, s1), ref)


def test_tiling_and_census_restatements_match_the_reference_dataset_methods(tmp_path):
    """the oracle's restated tile grid, centre mask, census aggregation and dasymetric
    adjustment against the reference's own Population_Dataset methods (data/PopulationDataset.py:294-334, 656-672, 675-729,
    823-852), called unbound on a bare instance with a stub rasterio that serves the synthetic boundary raster."""
    import numpy as np
    H, W, R = 300, 260, 12
    ids = po.synthetic_regions(H, W, R)
    bboxes = po.region_bboxes(ids, R)
    keep = [r for r in range(1, R + 1) if bboxes[r - 1] is not None]
    pop = [1000.0 * r + 17 for r in keep]
    csv = tmp_path / "census.csv"
    with open(csv, "w") as f:
        f.write("idx,bbox,POP20\n")
        for r, p in zip(keep, pop):
            f.write(f'{r},"{list(bboxes[r - 1])}",{p}\n')
    cls = rs.load_reference_dataset_class({"boundary.tif": ids.numpy().astype(np.int32)})
    ds = object.__new__(cls)
    ds.file_paths = {"fine": {"boundary": "boundary.tif", "census": str(csv)}}
    ds.train_level = "fine"
    ds.fourseasons = False

    for (h, w, ps, ov) in ((300, 260, 96, 16), (97, 400, 96, 16), (2500, 2100, 2048, 128), (96, 96, 96, 16)):
        ds.img_shape = (h, w)
        ref = ds.get_patch_indices(ps, ov)
        assert torch.equal(ref[:, :2], po.get_patch_indices(h, w, ps, ov)) and bool((ref[:, 2] == 0).all())
        assert torch.equal(torch.from_numpy(ds._create_mask(ps, ps + 32, ov)), po.centre_mask(ps, ps + 32, ov))

    g = torch.Generator().manual_seed(3)
    pred = torch.rand(H, W, generator=g) * (torch.rand(H, W, generator=g) > 0.3)
    pred[ids == keep[0]] = 0.0                       # a region whose predicted sum is 0 is left unscaled (:846-847)
    ref_pred, ref_census = ds.convert_popmap_to_census(pred.clone(), gpu_mode=False)
    mine = po.convert_popmap_to_census(pred, ids.float(), keep, [bboxes[r - 1] for r in keep])
    assert torch.equal(ref_pred, mine[mine > -1]) and torch.equal(ref_census, torch.tensor(pop))
    # the one-pass segment sum the CUDA path implements gives the same totals
    assert torch.allclose(po.region_sums(pred, ids, R + 1)[keep].float(), ref_pred, rtol=1e-5)
    ref_adj = ds.adjust_map_to_census(pred.clone())
    assert torch.equal(ref_adj, po.adjust_map_to_census(pred, ids.float(), keep, [bboxes[r - 1] for r in keep], pop))


def test_restated_eval_loop_matches_the_reference_trainer_loop(tmp_path):
    """oracle.tiled_eval against the reference's OWN evaluation loop — run_eval.Trainer.test_target
    (run_eval.py:83-203) driven on CPU at a small tile size with two real reference models: accumulate, visit counts,
    mean / std of the density and scale maps.  Both loops call the same reference forward, so the maps must agree bit for bit."""
    import argparse
    import json
    import os
    import types
    import fake_device as fd
    re_mod = rs.load_reference_run_eval()
    ps, ov, H, W = 96, 16, 200, 236
    members = [rs.build_reference_model(seed=1601).eval(), rs.build_reference_model(seed=1602).eval()]
    s2_file, s1 = po.synthetic_raw(H, W, seed=11)
    s2_rgbn = s2_file[[2, 1, 0, 3]].float()
    saved, to_census = {}, []

    class DS:
        region = "uga"                                       # testlevels_eval["uga"] = ["coarse"]: one aggregation pass

        def shape(self):
            return H, W

        def convert_popmap_to_census(self, m, gpu_mode=False, level=None, details_to=None):
            to_census.append(m.clone())
            return torch.tensor([1.0, 2.0]), torch.tensor([1.0, 3.0])

        def adjust_map_to_census(self, m):
            return m

        def save(self, m, folder, tag=""):
            saved[tag] = m.clone()

    class Loader:
        dataset = DS()

        def __len__(self):
            return len(po.get_patch_indices(H, W, ps, ov))

        def __iter__(self):
            mask = po.centre_mask(ps, ps, ov)[None]
            for xl, yl in po.get_patch_indices(H, W, ps, ov).tolist():    # pinned to the reference's grid in the test above
                yield {"S2": s2_rgbn[None, :, xl:xl + ps, yl:yl + ps].clone(), "S1": s1[None, :, xl:xl + ps, yl:yl + ps].clone(),
                       "img_coords": [torch.tensor([xl]), torch.tensor([yl])], "mask": mask.clone()}

    with open(os.path.join(rs.REF_ROOT, "data", "config", "dataset_stats.json")) as f:
        stats = json.load(f)
    for k in stats:
        for kk in stats[k]:
            stats[k][kk] = torch.tensor(stats[k][kk])
    tr = object.__new__(re_mod.Trainer)
    tr.model, tr.dataset_stats, tr.dataloaders = members, stats, {"test_target": [Loader()]}
    tr.args = argparse.Namespace(buildinginput=False, segmentationinput=False)
    tr.info, tr.experiment_folder = {"iter": 0}, str(tmp_path)
    orig = (re_mod.torch, re_mod.ips, re_mod.wandb, torch.Tensor.cuda)
    re_mod.torch, re_mod.ips = fd._TorchProxy(), ps
    re_mod.wandb = types.SimpleNamespace(log=lambda *a, **k: None)
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        with rs.reference_cwd():
            tr.test_target(save=True)
    finally:
        re_mod.torch, re_mod.ips, re_mod.wandb, torch.Tensor.cuda = orig

    norm = po.read_and_normalize(s2_file, s1)[0]
    fwd = lambda m, inp: rs.reference_forward(m, inp, padding=False)
    with torch.no_grad():
        out, std, scale, count, scale_std = po.tiled_eval(members, norm, ps, ov, forward_fn=fwd, with_scale_std=True)
    eq = lambda a, b: torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0))
    assert eq(saved[""], out) and eq(to_census[0], out)
    assert eq(saved["STD"], std)
    assert eq(saved["SCALE_uga"], scale)
    assert eq(saved["SCALE_STD"], scale_std)
    assert int(count[ov:-ov, ov:-ov].min()) == 2 and int(count.max()) == 8       # 2 members x (1 .. 4 covering tiles)
    assert float(count[:ov].sum() + count[:, :ov].sum()) == 0.0                 # the outer frame is never written


def test_census_loss_matches_the_reference_get_loss():
    """The training objective used by the train-step parity tests and bench.py (oracle.train_loss == popcorn_b200.synthetic.
    census_loss) against utils/losses.py:get_loss with the README's training flags (loss=["log_l1_loss"], lam=[1.0],
    scale_regularization=0.01) times lam_weak=100 (run_train.py:205-213; arguments/train.py defaults), values and gradients."""
    from popcorn_b200 import synthetic as sy
    rs.load_reference()
    import utils.losses as rl
    g = torch.Generator().manual_seed(0)
    for n_scale in (0, 37):
        pop = (torch.rand(3, generator=g) * 5e3).requires_grad_(True)
        scale = (torch.randn(n_scale, generator=g)).requires_grad_(True) if n_scale else None
        y = torch.rand(3, generator=g) * 4e4
        out = {"popcount": pop, "popdensemap": torch.zeros(3, 4, 4), "scale": scale}
        ref, _ = rl.get_loss(out, {"y": y}, scale=scale, loss=["log_l1_loss"], lam=[1.0], scale_regularization=0.01, tag="weak")
        ref = ref * 100.0
        gref = torch.autograd.grad(ref, [pop] + ([scale] if n_scale else []))
        for fn in (po.train_loss, sy.census_loss):
            mine = fn(out, y)
            assert torch.allclose(mine, ref, rtol=1e-6)
            gm = torch.autograd.grad(mine, [pop] + ([scale] if n_scale else []))
            assert all(torch.allclose(a, b, rtol=1e-6, atol=1e-12) for a, b in zip(gm, gref))
