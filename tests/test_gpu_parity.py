"""GPU: parity of the CUDA path (through the reference-shaped model API and the C-ABI) against the golden
vectors produced by the unmodified reference, and against the CPU oracle on fresh inputs.

Bars (BASELINE.json north_star): sparse index set bit-exact; per-pixel density within 1e-2 relative
(|a-b| / max(|b|, 1e-3*max|b|)); region counts within 1e-3 relative.
"""
import pytest
import torch

from popcorn_b200 import ops, weights
from oracle import popcorn_oracle as po
from util import TOL_GRAD, TOL_PIXEL, TOL_REGION, build_model, golden, golden_state_dict, max_rel

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


def _kernel_vs_autograd(m, inp, y):
    import __graft_entry__ as ge
    return ge._head_backward_vs_autograd(m, inp, y)


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
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-3 * abs(float(g["loss"]))
    loss.backward()
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert sorted(grads) == sorted(k[5:] for k in g if k.startswith("grad."))   # only head.* receive gradients
    # gradient scale = the loss terms' own gradients (po.train_loss_terms: the per-region terms may cancel)
    total, per, _ = po.head_grad_terms(sd, {"input": g["input"], "admin_mask": g["admin_mask"], "census_idx": g["census_idx"]},
                                       g["y"], grid=(g["grid_x"], g["grid_y"]), padding=False)
    for k in grads:                                                              # the oracle reproduces the reference's gradients
        assert max_rel(total[k], g["grad." + k], floor_frac=1e-2) < 1e-4, k
    e_norm, e_elem = po.grad_parity_errors(grads, {k: g["grad." + k] for k in grads}, per)
    assert e_norm < TOL_GRAD and e_elem < 2 * TOL_GRAD, (e_norm, e_elem)
    assert _kernel_vs_autograd(m, inp, g["y"].cuda()) < 1e-4


# (name, boxes of the two regions in a 96x128 batch, census targets): n % 128 = 96 is __graft_entry__.smoke()'s batch
# (region 0 over-, region 1 under-predicted on the random weights: the two log-L1 gradients nearly cancel), 0 = no tail
# tile, 36 = the golden's tail length
_TRAIN_CASES = [("tail96_mixed_sign", [(10, 70, 20, 100), (30, 90, 8, 64)], [2500.0, 9000.0], 96),
                ("tail0", [(0, 64, 0, 64), (0, 32, 0, 128)], [2500.0, 9000.0], 0),
                ("tail36", [(10, 71, 20, 104), (30, 90, 8, 64)], [3500.0, 12000.0], 36)]


@pytest.mark.parametrize("weights", ["golden", "random1600"])
@pytest.mark.parametrize("case", _TRAIN_CASES, ids=[c[0] for c in _TRAIN_CASES])
def test_train_step_scale_and_gradients_vs_oracle(sd, weights, case):
    """Census train step (run_train.py:201-230, unet_no_grad) on the parity weights AND the benchmark weights: index set
    bit-exact, scale values, popcount, loss, and all 8 head gradients against the oracle; the backward kernel alone
    against torch.float64 autograd over the same features."""
    _, boxes, ys, tail = case
    w = sd if weights == "golden" else po.random_state_dict(seed=1600)
    B, H, W = 2, 96, 128
    x = po.synthetic_input(H, W, seed=3, B=B)
    admin = torch.zeros(B, H, W)
    for b, (r0, r1, c0, c1) in enumerate(boxes):
        admin[b, r0:r1, c0:c1] = float(4 + 5 * b)
    cidx, y = torch.tensor([4, 9]), torch.tensor(ys)
    m = build_model(w).train()
    torch.manual_seed(7)
    grid = po.sparsity_grid(H, W)
    torch.manual_seed(7)
    inp = {"input": x.cuda(), "admin_mask": admin.cuda(), "census_idx": cidx.cuda()}
    out = m(inp, train=True, padding=False, encoder_no_grad=True, unet_no_grad=True, sparse=True)
    loss = po.train_loss(out, y.cuda())
    loss.backward()
    total, per, ref = po.head_grad_terms(w, {"input": x, "admin_mask": admin, "census_idx": cidx}, y, grid=grid, padding=False)
    idx, n = m._last_compaction
    n = int(n.item())
    assert n % 128 == tail and n == int(ref["mask"].sum())
    mask = torch.zeros(B * H * W, dtype=torch.bool)
    mask[idx[:n].long().cpu()] = True
    assert torch.equal(mask.view(B, H, W), ref["mask"])
    assert out["scale"].shape == ref["scale"].shape
    grads = {k: p.grad for k, p in m.named_parameters() if k.startswith("head.")}
    e_norm, e_elem = po.grad_parity_errors(grads, total, per)
    got = {"scale": max_rel(out["scale"], ref["scale"]),              # values, in compaction (row-major) order
           "dens": max_rel(out["popdensemap"], ref["popdensemap"]),
           "popcount": max_rel(out["popcount"], ref["popcount"], floor_frac=1.0),
           "loss": abs(float(loss.detach()) - float(po.train_loss(ref, y).detach())) / abs(float(loss.detach())),
           "grad_norm": e_norm, "grad_elem": e_elem, "kernel": _kernel_vs_autograd(m, inp, y.cuda())}
    # bars: BASELINE.json's (1e-2 per pixel, 1e-3 per region) on the parity weights, whose activations reach 1e2-1e4 with heavy
    # cancellation (profiles/r1c_precision_study.md); the benchmark weights are three orders better conditioned and held to 1e-4 / 1e-5
    bars = {"scale": TOL_PIXEL, "dens": TOL_PIXEL, "popcount": TOL_REGION, "loss": 1e-3, "grad_norm": TOL_GRAD, "grad_elem": 2 * TOL_GRAD,
            "kernel": 1e-4}
    if weights != "golden":
        bars.update(scale=1e-4, dens=1e-4, popcount=1e-5, loss=1e-5)
    assert all(got[k] < bars[k] for k in bars), (got, bars)


def test_train_step_config3_size_gradients(sd):
    """BASELINE config 3's batch shape (B=2 x 896x960, ~1e6 selected pixels): with this many pixels single ReLU-knee flips
    are negligible and the plain element-wise relative error (floor 1e-2 of the largest element) meets 1e-3 as well."""
    B, H, W = 2, 896, 960
    x = po.synthetic_input(H, W, seed=11, B=B)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    admin = torch.zeros(B, H, W)
    admin[0][((yy - 448) / 400.0) ** 2 + ((xx - 480) / 420.0) ** 2 < 1] = 17.0
    admin[1][((yy - 430) / 380.0) ** 2 + ((xx - 500) / 400.0) ** 2 < 1] = 5.0
    cidx = torch.tensor([17, 5])
    m = build_model(sd).train()
    torch.manual_seed(7)
    grid = po.sparsity_grid(H, W)
    with torch.no_grad():
        ref0 = po.forward(sd, {"input": x, "admin_mask": admin, "census_idx": cidx}, padding=False, sparse=True, grid=grid)
    y = (ref0["popcount"] * torch.tensor([1.6, 2.3])).float()        # both regions under-predicted: no cancellation
    torch.manual_seed(7)
    inp = {"input": x.cuda(), "admin_mask": admin.cuda(), "census_idx": cidx.cuda()}
    out = m(inp, train=True, padding=False, encoder_no_grad=True, unet_no_grad=True, sparse=True)
    po.train_loss(out, y.cuda()).backward()
    total, per, ref = po.head_grad_terms(sd, {"input": x, "admin_mask": admin, "census_idx": cidx}, y, grid=grid, padding=False)
    assert max_rel(out["popcount"], ref["popcount"], floor_frac=1.0) < 1e-5
    grads = {k: p.grad for k, p in m.named_parameters() if k.startswith("head.")}
    e_norm, e_elem = po.grad_parity_errors(grads, total, per)
    assert e_norm < TOL_GRAD and e_elem < TOL_GRAD, (e_norm, e_elem)
    for k, v in grads.items():
        assert max_rel(v, total[k], floor_frac=1e-2) < TOL_GRAD, k


def test_smoke_entry_point():
    """__graft_entry__.smoke() is what the driver runs on a fresh B200 at round end: keep it green in every GPU test run."""
    import __graft_entry__ as ge
    ge.smoke()


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


def test_bench_sized_window_3840x16384_crops_vs_oracle(model, sd):
    """One merged window of the size bench.py runs (3840 x 16384: TMA maps and index arithmetic beyond 2^24 pixels, 30 column tiles
    of 128 x 60 row tiles per job): interior crops — first / last column tiles, rows beyond the first tile, the far corner — against
    the oracle run on the crop plus a 64-px frame (origins multiples of 4 keep the pool phase)."""
    H, W = 3840, 16384
    g = torch.Generator(device="cuda").manual_seed(1610)
    x = torch.empty(1, 6, H, W, device="cuda")
    for c in range(6):
        low = torch.nn.functional.interpolate(torch.randn(1, 1, H // 64 + 2, W // 64 + 2, generator=g, device="cuda"), size=(H, W),
                                              mode="bilinear", align_corners=True)[0, 0]
        x[0, c] = 0.6 * low + 0.8 * torch.randn(H, W, generator=g, device="cuda")
        del low
    with torch.no_grad():
        bu = ops.dda_forward(model._dda_pack("building_extractor"), x, model.p2d, ops.PC_DDA_BUILTUP)
        feats = ops.dda_forward(model._dda_pack("unetmodel"), x, (0, 0, 0, 0), ops.PC_DDA_FEATURES)
        dens, scale = ops.head_dense_forward(model._head_pack(tc=True), feats, bu, None, None, None, want_scale=True, tc=True)
        del feats
        F_, C = 64, 192
        worst = 0.0
        for y0, x0 in [(128, 128), (2048, 8192 - 64), (H - C - 2 * F_ - 4, W - C - 2 * F_ - 4), (1792, 16384 - 512), (3000, 4)]:
            y0, x0 = y0 // 4 * 4, x0 // 4 * 4
            crop = x[:, :, y0:y0 + C + 2 * F_, x0:x0 + C + 2 * F_].cpu()
            ref = po.forward(sd, {"input": crop}, padding=False)
            a = dens[0, y0 + F_:y0 + F_ + C, x0 + F_:x0 + F_ + C].cpu()
            b = ref["popdensemap"][0][F_:-F_, F_:-F_]
            worst = max(worst, max_rel(a, b))
            assert max_rel(bu[0, 0, y0 + F_:y0 + F_ + C, x0 + F_:x0 + F_ + C].cpu(), ref["popdensemap"][0][F_:-F_, F_:-F_] * 0 +
                           po.building_score(sd, crop)[0, 0][F_:-F_, F_:-F_]) < 1e-3
        assert worst < TOL_PIXEL, worst
        assert torch.equal(dens, scale * bu[:, 0])


@pytest.mark.parametrize("shape", [(1, 64, 96), (2, 75, 101), (1, 130, 70)])
def test_fused_eval_entry_point_equals_the_three_separate_calls(model, shape):
    """pc_infer_tile_fused (builtup pass + feature pass + head + census partials in one C-ABI call) against the same
    work issued as pc_dda_forward x 2 + pc_head_dense_forward_tc from Python: bit-identical outputs, with and without census ids."""
    from popcorn_b200.model import popcorn as pm
    B, H, W = shape
    x = po.synthetic_input(H, W, seed=H + W, B=B).cuda()
    admin = (torch.arange(B * H * W).view(B, H, W) % 3).float().cuda()
    cidx = torch.tensor([1] * B).cuda()
    outs = []
    for fused in (True, False):
        pm.FUSED_EVAL = fused
        try:
            with torch.no_grad():
                a = model({"input": x}, padding=False)
                b = model({"input": x, "admin_mask": admin, "census_idx": cidx}, padding=False)
        finally:
            pm.FUSED_EVAL = True
        outs.append((a, b))
    for k in ("popdensemap", "scale", "builtup_score", "popcount"):
        assert torch.equal(outs[0][0][k], outs[1][0][k]), k
        assert torch.equal(outs[0][1][k], outs[1][1][k]), k
    ref = po.forward(golden_state_dict(), {"input": x.cpu(), "admin_mask": admin.cpu(), "census_idx": cidx.cpu()}, padding=False)
    assert max_rel(outs[0][1]["popcount"], ref["popcount"], floor_frac=1.0) < TOL_REGION
