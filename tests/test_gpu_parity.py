"""GPU: parity of the CUDA path (through the reference-shaped model API and the C-ABI) against the golden
vectors produced by the unmodified reference, and against the CPU oracle on fresh inputs.

Bars (BASELINE.json north_star): sparse index set bit-exact; per-pixel density within 1e-2 relative
(|a-b| / max(|b|, 1e-3*max|b|)); region counts within 1e-3 relative.
"""
import pytest
import torch

from popcorn_b200 import ops, weights
from oracle import popcorn_oracle as po
from util import TOL_PIXEL, TOL_REGION, build_model, golden, golden_state_dict, max_rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sd():
    return golden_state_dict()


@pytest.fixture(scope="module")
def model(sd):
    return build_model(sd).eval()


@pytest.mark.parametrize("name", ["dense_64x96", "dense_75x101", "dense_pad14_48x80", "dense_130x70"])
def test_dense_forward_vs_reference_golden(model, name):
    g = golden(name)
    inp = {"input": g["input"].cuda()}
    with torch.no_grad():
        out = model(inp, padding=bool(g["padding"]))
    assert set(("popcount", "popdensemap", "scale")) <= set(out)
    assert out["popdensemap"].shape == g["popdensemap"].shape
    assert max_rel(inp["building_counts"], g["builtup"]) < 1e-4
    assert max_rel(out["scale"], g["scale"]) < TOL_PIXEL
    assert max_rel(out["popdensemap"], g["popdensemap"]) < TOL_PIXEL
    assert max_rel(out["popcount"], g["popcount"], floor_frac=1.0) < TOL_REGION


@pytest.mark.parametrize("copy,mode", [("unetmodel", 0), ("building_extractor", 1)])
@pytest.mark.parametrize("H,W", [(64, 64), (97, 131), (33, 250), (256, 192), (150, 530)])
@pytest.mark.parametrize("tc", [False, True], ids=["simt", "tcgen05"])
def test_dda_forward_vs_oracle(sd, copy, mode, H, W, tc):
    x = po.synthetic_input(H, W, seed=H * W)
    pack = weights.pack_dda(sd, copy, tc=tc).cuda()
    if mode == 0:
        pads = po.feature_padding(H, W, False)
        ref = po.unet_features(sd, x, padding=False)
    else:
        pads = (14, 14, 14, 14)
        ref = po.building_score(sd, x)
    got = ops.dda_forward(pack, x.cuda(), pads, mode)
    assert got.shape == ref.shape
    assert max_rel(got, ref, floor_frac=1.0) < 2e-5          # fp32 rounding relative to the largest activation
    assert max_rel(got, ref) < 5e-3                            # per element, floor 1e-3 * max


@pytest.mark.parametrize("C", [2, 4])
def test_single_modality_variants(sd, C):
    """input_channels 2 (S1 only) / 4 (S2+NIR only): only the present stream runs, 8-ch head input (popcorn.py:48-54)."""
    H, W = 64, 96
    x6 = po.synthetic_input(H, W, seed=C)
    x = x6[:, 4:6] if C == 2 else x6[:, :4]
    sdc = dict(sd)
    g = torch.Generator().manual_seed(C)
    sdc["head.0.weight"] = (torch.rand(64, 8, 1, 1, generator=g) - 0.5) * 0.7
    with torch.no_grad():
        ref = po.forward(sdc, {"input": x.clone()}, padding=False)
    m = build_model(sdc, input_channels=C).eval()
    with torch.no_grad():
        out = m({"input": x.cuda()}, padding=False)
    assert max_rel(out["popdensemap"], ref["popdensemap"]) < TOL_PIXEL


def test_sparse_train_step_vs_reference_golden(sd):
    g = golden("sparse_train")
    m = build_model(sd).train()
    inp = {"input": g["input"].cuda(), "admin_mask": g["admin_mask"].cuda(), "census_idx": g["census_idx"].cuda()}
    H, W = g["input"].shape[2:]
    # the golden grid came from torch.manual_seed(4242) followed by the two multinomial draws (popcorn.py:367-368)
    torch.manual_seed(4242)
    out = m(inp, train=True, padding=False, encoder_no_grad=True, unet_no_grad=True, sparse=True)
    mask = torch.zeros(g["mask"].numel(), dtype=torch.bool)
    idx, n = m._last_compaction
    mask[idx[:int(n.item())].long().cpu()] = True
    assert torch.equal(mask.view_as(g["mask"]), g["mask"])                        # bit-exact index set
    assert out["scale"].shape == g["scale"].shape
    assert max_rel(out["scale"], g["scale"]) < TOL_PIXEL
    assert max_rel(out["popdensemap"], g["popdensemap"]) < TOL_PIXEL
    assert max_rel(out["popcount"], g["popcount"], floor_frac=1.0) < TOL_REGION
    loss = po.train_loss(out, g["y"].cuda())
    assert abs(float(loss) - float(g["loss"])) < 1e-3 * abs(float(g["loss"]))
    loss.backward()
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert sorted(grads) == sorted(k[5:] for k in g if k.startswith("grad."))   # only head.* receive gradients
    for k, v in grads.items():
        assert max_rel(v, g["grad." + k], floor_frac=1e-2) < 1e-2, k


def test_backward_is_deterministic(sd):
    g = golden("sparse_train")
    m = build_model(sd).train()
    res = []
    for _ in range(2):
        for p in m.parameters():
            p.grad = None
        torch.manual_seed(4242)
        inp = {"input": g["input"].cuda(), "admin_mask": g["admin_mask"].cuda(), "census_idx": g["census_idx"].cuda()}
        out = m(inp, train=True, padding=False, unet_no_grad=True, sparse=True)
        po.train_loss(out, g["y"].cuda()).backward()
        res.append(torch.cat([p.grad.reshape(-1) for p in m.head.parameters()]).clone())
    assert torch.equal(res[0], res[1])


def test_dense_head_with_admin_mask_popcount(model, sd):
    """Eval-mode call with admin_mask / census_idx (run_train.py validation): popcount = masked sum (:186-187)."""
    g = golden("sparse_train")
    inp = {"input": g["input"].cuda(), "admin_mask": g["admin_mask"].cuda(), "census_idx": g["census_idx"].cuda()}
    with torch.no_grad():
        out = model(inp, padding=False)
        ref = po.forward(sd, {"input": g["input"], "admin_mask": g["admin_mask"], "census_idx": g["census_idx"]},
                         padding=False)
    assert max_rel(out["popdensemap"], ref["popdensemap"]) < TOL_PIXEL
    assert max_rel(out["popcount"], ref["popcount"], floor_frac=1.0) < TOL_REGION


def test_full_tile_properties_2048(model):
    """BASELINE config 1 size (2048^2): size-independent properties instead of a CPU comparison."""
    H = W = 2048
    x = po.synthetic_input(H, W, seed=1610).cuda()
    ids = po.synthetic_regions(H, W, R=400).cuda()
    with torch.no_grad():
        out = model({"input": x}, padding=False)
        dens = out["popdensemap"][0]
        # (1) region sums partition the map: sum of sums == popcount == direct sum
        sums = ops.region_sum(dens.contiguous(), ids, 401)
        total = float(dens.double().sum())
        assert abs(float(sums.sum()) - total) < 1e-6 * total
        assert abs(float(out["popcount"][0]) - total) < 1e-4 * total
        # (2) translation equivariance for shifts that are multiples of 4: interior pixels agree
        out2 = model({"input": x[:, :, 64:, 128:].contiguous()}, padding=False)
        a = dens[64 + 40:-40, 128 + 40:-40]
        b = out2["popdensemap"][0][40:-40, 40:-40]
        assert max_rel(b, a) < 1e-4
        # (3) density = occupancy x builtup, occupancy >= 0, builtup in (0,1)
        assert torch.equal(dens, out["scale"][0] * out["builtup_score"][0, 0])
        assert float(out["scale"].min()) >= 0 and 0 < float(out["builtup_score"].min()) and float(out["builtup_score"].max()) < 1
    # (4) against the oracle on a 256x256 interior crop re-run as its own tile origin-aligned to 4
    sd = golden_state_dict()
    crop = x[:, :, 512:768, 1024:1280].cpu()
    with torch.no_grad():
        ref = po.forward(sd, {"input": crop}, padding=False)["popdensemap"][0]
    assert max_rel(dens[512 + 40:768 - 40, 1024 + 40:1280 - 40].cpu(), ref[40:-40, 40:-40]) < TOL_PIXEL
