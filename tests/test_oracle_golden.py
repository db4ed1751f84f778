"""CPU: the oracle restatement reproduces the committed golden vectors, which were produced by the
imported, unmodified reference (oracle/make_golden.py).  This is the parity pin of the oracle."""
import pytest
import torch

from oracle import popcorn_oracle as po
from util import golden, golden_state_dict, max_rel

torch.set_num_threads(max(1, min(8, torch.get_num_threads())))


@pytest.fixture(scope="module")
def sd():
    return golden_state_dict()


@pytest.mark.parametrize("name", ["dense_64x96", "dense_75x101", "dense_pad14_48x80", "dense_130x70"])
def test_dense_forward_matches_reference_golden(sd, name):
    g = golden(name)
    inp = {"input": g["input"].clone()}
    with torch.no_grad():
        out = po.forward(sd, inp, padding=bool(g["padding"]))
    assert max_rel(inp["building_counts"], g["builtup"]) < 1e-3
    assert max_rel(out["popdensemap"], g["popdensemap"]) < 1e-3
    assert max_rel(out["scale"], g["scale"]) < 1e-3
    assert max_rel(out["popcount"], g["popcount"]) < 1e-3


def test_sparse_train_step_matches_reference_golden(sd):
    g = golden("sparse_train")
    sdg = {k: (v.clone().requires_grad_(True) if k.startswith("head.") else v) for k, v in sd.items()}
    inp = {"input": g["input"].clone(), "admin_mask": g["admin_mask"].clone(), "census_idx": g["census_idx"].clone()}
    out = po.forward(sdg, inp, padding=False, sparse=True, grid=(g["grid_x"], g["grid_y"]))
    assert torch.equal(out["mask"], g["mask"])                       # index set: bit-exact
    assert out["scale"].shape == g["scale"].shape
    assert max_rel(out["scale"], g["scale"]) < 1e-3
    assert max_rel(out["popcount"], g["popcount"]) < 1e-3
    loss = po.train_loss(out, g["y"])
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    loss.backward()
    for k in [k for k in g if k.startswith("grad.")]:
        assert max_rel(sdg[k[5:]].grad, g[k], floor_frac=1e-2) < 1e-3, k


def test_loss_terms_sum_to_the_loss_and_their_gradients_to_the_reference_gradient(sd):
    """po.train_loss_terms / head_grad_terms (the gradient-scale helpers of the GPU parity tests): the terms add up to
    train_loss and their gradients to the gradients the imported reference produced (tests/golden/sparse_train.npz)."""
    g = golden("sparse_train")
    inp = {"input": g["input"].clone(), "admin_mask": g["admin_mask"].clone(), "census_idx": g["census_idx"].clone()}
    total, per, out = po.head_grad_terms(sd, inp, g["y"], grid=(g["grid_x"], g["grid_y"]), padding=False)
    assert len(per) == g["y"].numel() + 1
    terms = po.train_loss_terms(out, g["y"])
    assert abs(float(sum(terms)) - float(po.train_loss(out, g["y"]))) < 1e-5 * abs(float(g["loss"]))
    for k in [k for k in g if k.startswith("grad.")]:
        assert max_rel(total[k[5:]], g[k], floor_frac=1e-2) < 1e-3, k
    e_norm, e_elem = po.grad_parity_errors({k[5:]: g[k] for k in g if k.startswith("grad.")}, total, per)
    assert e_norm < 1e-5 and e_elem < 1e-5


def test_tiled_eval_matches_reference_golden(sd):
    g = golden("tiled_eval")
    ps, ov = int(g["patchsize"]), int(g["overlap"])
    with torch.no_grad():
        m, _, smap, cnt = po.tiled_eval([sd], g["raster"], ps, ov)
    assert torch.equal(cnt, g["count"])
    assert max_rel(m, g["map"]) < 1e-3
    assert max_rel(smap, g["scale_map"]) < 1e-3
    bboxes = [tuple(b.tolist()) if b[0] >= 0 else None for b in g["bboxes"]]
    census = po.convert_popmap_to_census(m, g["ids"].float(), list(range(1, len(bboxes) + 1)), bboxes)
    assert max_rel(census, g["census"]) < 1e-3
    # one-pass region sums == the reference's per-region bbox loop
    rs = po.region_sums(m, g["ids"], len(bboxes) + 1)[1:]
    valid = census > -1
    assert max_rel(rs[valid].float(), census[valid]) < 1e-3


def test_patch_indices_cover_like_reference():
    # data/PopulationDataset.py:294-316 — main grid + bottom row + right column + corner
    idx = po.get_patch_indices(5000, 4500, 2048, 128)
    assert idx[0].tolist() == [0, 0]
    assert [5000 - 2048, 4500 - 2048] in idx.tolist()
    xs = sorted(set(idx[:, 0].tolist()))
    assert xs == [0, 1792, 5000 - 2048]


def test_random_state_dict_has_reference_keys(sd):
    r = po.random_state_dict()
    assert sorted(r.keys()) == sorted(sd.keys())
    for k in sd:
        assert tuple(r[k].shape) == tuple(sd[k].shape), k


@pytest.mark.parametrize("enc", [False, True])
def test_finetune_gradients_match_reference_golden(sd, enc):
    """N4 pin: torch.autograd through the oracle == the gradients the imported reference produced for the fine-tuning
    step (unet_no_grad=False, encoder_no_grad=enc; oracle/make_golden_finetune.py): same set of tensors (conv / convT
    weights and biases of unetmodel + head; BN frozen; with encoder_no_grad the encoder gets none) and same values."""
    g = golden("finetune")
    tag = "enc1" if enc else "enc0"
    gkeys = sorted(k[len(tag) + 6:] for k in g if k.startswith(tag + ".grad."))
    assert len(gkeys) == (32 if enc else 56)
    x, admin, cidx, y = g["input"], g["admin_mask"], g["census_idx"], g["y"]
    is_param = lambda k, v: (k.startswith("head.") or k.startswith("unetmodel.")) and v.is_floating_point() and "running_" not in k
    sdg = {k: (v.clone().requires_grad_(True) if is_param(k, v) else v) for k, v in sd.items()}
    out = po.forward(sdg, {"input": x, "admin_mask": admin, "census_idx": cidx}, padding=False, sparse=True,
                     grid=(g[tag + ".grid_x"], g[tag + ".grid_y"]), encoder_no_grad=enc)
    loss = po.train_loss(out, y)
    loss.backward()
    assert abs(float(loss) - float(g[tag + ".loss"])) < 1e-5 * abs(float(g[tag + ".loss"]))
    for k in gkeys:
        ref = g[f"{tag}.grad.{k}"]
        got = sdg[k].grad
        assert got is not None, k
        assert float((got - ref).abs().max()) <= 1e-4 * float(ref.abs().max()) + 1e-12, k
    # nothing else of unetmodel receives a gradient in the reference (BN affine frozen; encoder under no_grad if enc)
    extra = [k for k, v in sdg.items() if k.startswith("unetmodel.") and v.is_floating_point() and v.requires_grad
             and v.grad is not None and float(v.grad.abs().max()) > 0 and k not in gkeys]
    bn_or_stats = lambda k: k.split(".")[-2] in ("1", "4") or "out" in k
    assert all(bn_or_stats(k) for k in extra), [k for k in extra if not bn_or_stats(k)]
