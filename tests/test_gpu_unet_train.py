"""GPU: fine-tuning path — POPCORN.forward with unet_no_grad=False back-propagates into the conv /
transposed-conv parameters of `unetmodel` through the hand-written backward (csrc/unet_bwd.cu, model/unet_train.py);
gradients are compared with torch.autograd through the oracle's functional restatement of the reference
(run_train.py:191-238, model/DDA_model/utils/networks.py:121-151)."""
import pytest
import torch

from popcorn_b200.model import unet_train
from oracle import popcorn_oracle as po
from util import TOL_GRAD, TOL_REGION, build_model, golden_state_dict, max_rel

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _case(B, H, W, seed):
    x = po.synthetic_input(H, W, seed=seed, B=B)
    admin = torch.zeros(B, H, W)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    for b in range(B):
        admin[b][((yy - H / 2) / (0.4 * H)) ** 2 + ((xx - W / 2) / (0.35 * W)) ** 2 < 1] = float(3 + b)
    cidx = torch.tensor([3 + b for b in range(B)])
    y = torch.tensor([2500.0 * (b + 1) for b in range(B)])
    return x, admin, cidx, y


@pytest.mark.parametrize("shape", [(2, 64, 96), (1, 75, 101), (2, 40, 136)])
@pytest.mark.parametrize("encoder_no_grad", [False, True])
def test_unet_finetune_gradients_match_autograd_of_the_reference_restatement(shape, encoder_no_grad):
    B, H, W = shape
    sd = golden_state_dict()
    model = build_model(sd).train()
    x, admin, cidx, y = _case(B, H, W, seed=H + W)
    torch.manual_seed(11)
    grid = po.sparsity_grid(H, W)
    torch.manual_seed(11)
    out = model({"input": x.cuda(), "admin_mask": admin.cuda(), "census_idx": cidx.cuda()}, train=True, padding=False,
                encoder_no_grad=encoder_no_grad, unet_no_grad=False, sparse=True)
    po.train_loss(out, y.cuda()).backward()

    keys = ["unetmodel." + k for k in unet_train.trainable_keys()] + [k for k in sd if k.startswith("head.")]
    total, per, ref = po.grad_terms(sd, {"input": x, "admin_mask": admin, "census_idx": cidx}, y, keys, grid=grid, padding=False,
                                    encoder_no_grad=encoder_no_grad)

    assert max_rel(out["popcount"], ref["popcount"], floor_frac=1.0) < TOL_REGION
    params = dict(model.named_parameters())
    encoder = ("inc.", "down_seq.")
    checked, got = 0, {}
    for k in keys:
        g, r = params[k].grad, total[k]
        if encoder_no_grad and k.startswith("unetmodel.") and any(e in k for e in encoder):
            assert g is None or float(g.abs().max()) == 0.0, k
            assert float(r.abs().max()) == 0.0
            continue
        assert g is not None, k
        assert g.shape == r.shape, k
        got[k] = g
        checked += 1
    # errors against the gradient scale of the loss terms (oracle.grad_parity_errors; these batches of a few thousand pixels are
    # exposed to single ReLU-knee flips, hence the element-wise bar of 2x)
    e_norm, e_elem = po.grad_parity_errors(got, total, per)
    assert e_norm < TOL_GRAD and e_elem < 2 * TOL_GRAD, (e_norm, e_elem)
    assert checked >= (24 if encoder_no_grad else 48)
    # BN affine parameters are frozen by freeze_bn_layers (networks.py:184-189): no gradient
    assert all(p.grad is None for n, p in params.items() if n.startswith("unetmodel.") and n.split(".")[-2] in ("1", "4"))


def test_optimizer_step_changes_unet_and_next_forward_uses_new_weights():
    """Adam over the reference's parameter groups moves unetmodel; the packed inference weights are re-folded."""
    sd = golden_state_dict()
    model = build_model(sd).train()
    x, admin, cidx, y = _case(1, 64, 64, seed=2)
    inp = lambda: {"input": x.cuda(), "admin_mask": admin.cuda(), "census_idx": cidx.cuda()}
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-3)
    torch.manual_seed(3)
    out = model(inp(), train=True, padding=False, sparse=True)
    po.train_loss(out, y.cuda()).backward()
    before = model.unetmodel.get_parameter("sar_stream.inc.conv.conv.0.weight").detach().clone()
    opt.step()
    after = model.unetmodel.get_parameter("sar_stream.inc.conv.conv.0.weight").detach()
    assert not torch.equal(before, after)
    model.eval()
    with torch.no_grad():
        e = model(inp(), padding=False)
        sd2 = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        r = po.forward(sd2, {"input": x, "admin_mask": admin, "census_idx": cidx}, padding=False)
    assert max_rel(e["popdensemap"], r["popdensemap"]) < 1e-2
