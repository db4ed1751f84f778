"""TEST INFRASTRUCTURE: generates tests/golden/*.npz from the IMPORTED, UNMODIFIED reference
(/root/reference, CPU fp32) and reports how closely oracle/popcorn_oracle.py restates it.

Run here (the reference cannot travel to the GPU box):   python oracle/make_golden.py

Model = BASELINE config 1: POPCORN(6, occupancymodel=True, pretrained=False, biasinit=0.9407,
sentinelbuildings=True), torch.manual_seed(1600): random-init unetmodel convs / head, DDA
checkpoint weights for building_extractor (model/popcorn.py:57-97).
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


def main():
    torch.set_num_threads(8)
    torch.backends.mkldnn.enabled = True
    os.makedirs(GOLD, exist_ok=True)
    model = rs.build_reference_model(seed=1600).eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    np.savez_compressed(os.path.join(GOLD, "state_dict.npz"), **{k: v.numpy() for k, v in sd.items()})
    report = []

    # ---- dense eval cases (run_eval.py:109 call shape: model(sample, padding=False)) ---------------------
    for name, (H, W), padding in (("dense_64x96", (64, 96), False), ("dense_75x101", (75, 101), False),
                                  ("dense_pad14_48x80", (48, 80), True), ("dense_130x70", (130, 70), False)):
        x = po.synthetic_input(H, W, seed=1610 + H)
        inp = {"input": x.clone()}
        with torch.no_grad():
            ref = rs.reference_forward(model, inp, padding=padding)
        oin = {"input": x.clone()}
        with torch.no_grad():
            ora = po.forward(sd, oin, padding=padding)
        report.append((name, "popdensemap", rel(ora["popdensemap"], ref["popdensemap"])))
        report.append((name, "builtup", rel(oin["building_counts"], inp["building_counts"])))
        report.append((name, "popcount", rel(ora["popcount"], ref["popcount"])))
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), input=x.numpy(), padding=np.array(padding),
                            popdensemap=ref["popdensemap"].numpy(), scale=ref["scale"].numpy(),
                            popcount=ref["popcount"].numpy(), builtup=inp["building_counts"].numpy())

    # ---- sparse census-supervised step (run_train.py:201-230) --------------------------------------------
    B, H, W = 2, 72, 88
    x = po.synthetic_input(H, W, seed=77, B=B)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    admin = torch.zeros(B, H, W)
    admin[0][((yy - 30) / 25.0) ** 2 + ((xx - 40) / 30.0) ** 2 < 1] = 17.0
    admin[1][((yy - 40) / 28.0) ** 2 + ((xx - 50) / 22.0) ** 2 < 1] = 5.0
    admin[1, :, 80:] = -1.0                                   # collate padding value (PopulationDataset.py:898-918)
    census_idx = torch.tensor([17, 5])
    y = torch.tensor([3500.0, 12000.0])
    torch.manual_seed(4242)
    grid = po.sparsity_grid(H, W)                             # the CPU RNG draws the reference will make
    torch.manual_seed(4242)
    model.train()
    for p in model.parameters():
        p.grad = None
    inp = {"input": x.clone(), "admin_mask": admin.clone(), "census_idx": census_idx.clone()}
    ref = rs.reference_forward(model, inp, train=True, padding=False, encoder_no_grad=True, unet_no_grad=True,
                               sparse=True)
    loss = po.train_loss(ref, y)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    assert sorted(grads) == sorted(f"head.{i}.{t}" for i in (0, 2, 4, 6) for t in ("weight", "bias")), sorted(grads)
    model.eval()
    sdg = {k: (v.clone().requires_grad_(True) if k.startswith("head.") else v) for k, v in sd.items()}
    oin = {"input": x.clone(), "admin_mask": admin.clone(), "census_idx": census_idx.clone()}
    ora = po.forward(sdg, oin, padding=False, sparse=True, grid=grid)
    oloss = po.train_loss(ora, y)
    oloss.backward()
    assert torch.equal(ora["mask"], (inp["admin_mask"] == census_idx.view(-1, 1, 1)) & ora["mask"])
    report.append(("sparse_train", "popcount", rel(ora["popcount"], ref["popcount"])))
    report.append(("sparse_train", "scale", rel(ora["scale"], ref["scale"])))
    report.append(("sparse_train", "loss", rel(oloss.detach(), loss.detach())))
    for k in grads:
        report.append(("sparse_train", "grad " + k, rel(sdg[k].grad, grads[k])))
    # the mask the reference used = where its scattered output is defined: recover from scale length
    assert ref["scale"].numel() == int(ora["mask"].sum()), (ref["scale"].numel(), int(ora["mask"].sum()))
    np.savez_compressed(os.path.join(GOLD, "sparse_train.npz"), input=x.numpy(), admin_mask=admin.numpy(),
                        census_idx=census_idx.numpy(), y=y.numpy(), grid_x=grid[0].numpy(), grid_y=grid[1].numpy(),
                        mask=ora["mask"].numpy(), popcount=ref["popcount"].detach().numpy(),
                        popdensemap=ref["popdensemap"].detach().numpy(), scale=ref["scale"].detach().numpy(),
                        builtup=inp["building_counts"].numpy(), loss=loss.detach().numpy(),
                        **{"grad." + k: v.numpy() for k, v in grads.items()})

    # ---- tiled eval (restated run_eval loop around the reference forward), small patch size ----------------
    Hh, Ww, ps, ov = 266, 301, 128, 32   # edge tiles at 138 / 173: off the 4-px pool phase of the main grid
    raster = po.synthetic_input(Hh, Ww, seed=99)[0]
    with torch.no_grad():
        ref_map, ref_std, ref_scale, ref_cnt = po.tiled_eval(
            [model], raster, ps, ov, forward_fn=lambda m, i: rs.reference_forward(m, i, padding=False))
        ora_map, _, _, ora_cnt = po.tiled_eval([sd], raster, ps, ov)
    ids = po.synthetic_regions(Hh, Ww, R=12, seed=3)
    bboxes = po.region_bboxes(ids, 12)
    census = po.convert_popmap_to_census(ref_map, ids.float(), list(range(1, 13)), bboxes)
    report.append(("tiled_eval", "map", rel(ora_map, ref_map)))
    assert torch.equal(ora_cnt, ref_cnt)
    np.savez_compressed(os.path.join(GOLD, "tiled_eval.npz"), raster=raster.numpy(), patchsize=np.array(ps),
                        overlap=np.array(ov), map=ref_map.numpy(), scale_map=ref_scale.numpy(),
                        count=ref_cnt.numpy(), ids=ids.numpy(), census=census.numpy(),
                        bboxes=np.array([b if b is not None else (-1, -1, -1, -1) for b in bboxes]))

    print("oracle vs imported reference (max |diff| / max |ref|):")
    worst = 0.0
    for case, what, r in report:
        print(f"  {case:22s} {what:28s} {r:.3e}")
        worst = max(worst, r)
    print("worst:", worst)
    with open(os.path.join(GOLD, "ORACLE_PIN.txt"), "w") as f:
        f.write("oracle/popcorn_oracle.py vs imported reference (CPU fp32, torch %s); max|diff|/max|ref|\n" % torch.__version__)
        for case, what, r in report:
            f.write(f"{case} {what} {r:.3e}\n")
    assert worst < 1e-4, worst


if __name__ == "__main__":
    main()
