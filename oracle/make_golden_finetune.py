"""TEST INFRASTRUCTURE: golden gradients of the FINE-TUNING step from the IMPORTED, UNMODIFIED
reference (/root/reference, CPU fp32): model(sample, train=True, padding=False, encoder_no_grad=E, unet_no_grad=False,
sparse=True) + log-L1 loss + backward (run_train.py:191-230) -> tests/golden/finetune.npz.  Also reports how closely
torch.autograd through oracle/popcorn_oracle.py reproduces them (appended to tests/golden/ORACLE_PIN.txt).

Run here (the reference cannot travel to the GPU box):   python oracle/make_golden_finetune.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import popcorn_oracle as po          # noqa: E402
from oracle import reference_shim as rs          # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def case():
    B, H, W = 2, 56, 88                                        # 56 % 32 != 0 -> reflect padding to 64 rows (popcorn.py:247-256)
    x = po.synthetic_input(H, W, seed=311, B=B)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    admin = torch.zeros(B, H, W)
    admin[0][((yy - 26) / 20.0) ** 2 + ((xx - 40) / 30.0) ** 2 < 1] = 7.0
    admin[1][((yy - 30) / 22.0) ** 2 + ((xx - 50) / 26.0) ** 2 < 1] = 2.0
    return x, admin, torch.tensor([7, 2]), torch.tensor([1800.0, 6400.0])


def main():
    torch.set_num_threads(8)
    sd_file = np.load(os.path.join(GOLD, "state_dict.npz"))
    sd = {k: torch.from_numpy(sd_file[k]) for k in sd_file.files}
    model = rs.build_reference_model(seed=1600)
    model.load_state_dict(sd)                                  # the committed config-1 weights
    x, admin, cidx, y = case()
    out = {"input": x.numpy(), "admin_mask": admin.numpy(), "census_idx": cidx.numpy(), "y": y.numpy()}
    report = []
    for enc in (False, True):
        tag = "enc1" if enc else "enc0"
        model.train()
        for p in model.parameters():
            p.grad = None
            p.requires_grad_(True)
        torch.manual_seed(2024)
        grid = po.sparsity_grid(x.shape[2], x.shape[3])
        torch.manual_seed(2024)
        inp = {"input": x.clone(), "admin_mask": admin.clone(), "census_idx": cidx.clone()}
        ref = rs.reference_forward(model, inp, train=True, padding=False, encoder_no_grad=enc, unet_no_grad=False, sparse=True)
        loss = po.train_loss(ref, y)
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
        # what receives gradients: head + unetmodel convs / transposed convs (BN is frozen by freeze_bn_layers)
        assert not any(k.startswith("building_extractor.") for k in grads)
        assert not any(k.startswith("unetmodel.") and k.split(".")[-2] in ("1", "4") for k in grads), "BN affine must be frozen"
        keys = sorted(k for k in sd if k.startswith("head.") or k in grads)
        sdg = {k: (v.clone().requires_grad_(True) if k in keys else v) for k, v in sd.items()}
        ora = po.forward(sdg, {"input": x.clone(), "admin_mask": admin.clone(), "census_idx": cidx.clone()}, padding=False,
                         sparse=True, grid=grid, encoder_no_grad=enc)
        oloss = po.train_loss(ora, y)
        oloss.backward()
        report.append((f"finetune_{tag}", "loss", rel(oloss.detach(), loss.detach())))
        report.append((f"finetune_{tag}", "popcount", rel(ora["popcount"], ref["popcount"])))
        worst_k, worst = None, 0.0
        for k, g in grads.items():
            og = sdg[k].grad
            assert og is not None, k
            r = rel(og, g)
            if r > worst:
                worst_k, worst = k, r
        report.append((f"finetune_{tag}", f"worst grad ({len(grads)} tensors): {worst_k}", worst))
        out[f"{tag}.loss"] = loss.detach().numpy()
        out[f"{tag}.popcount"] = ref["popcount"].detach().numpy()
        out[f"{tag}.grid_x"], out[f"{tag}.grid_y"] = grid[0].numpy(), grid[1].numpy()
        for k, g in grads.items():
            out[f"{tag}.grad.{k}"] = g.numpy()
    np.savez_compressed(os.path.join(GOLD, "finetune.npz"), **out)
    with open(os.path.join(GOLD, "ORACLE_PIN.txt"), "a") as f:
        for c, what, r in report:
            line = f"{c} {what} {r:.3e}"
            print(line)
            f.write(line + "\n")
    assert max(r for _, _, r in report) < 1e-4


if __name__ == "__main__":
    main()
