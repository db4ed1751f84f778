"""Bisect of the head-gradient parity of the census train step (development tool; GPU).

Runs the exact configuration of __graft_entry__.smoke()'s sparse step and splits the comparison into stages:
  A  our features / builtup / scale_sel against the CPU oracle,
  B  pc_head_sparse_backward against torch autograd ON THE GPU over OUR features (isolates csrc/head_bwd.cu),
     per gradient term (g_pop only, g_scale only, both), with the tcgen05 and the SIMT forward,
  C  the end-to-end number smoke() asserts, per parameter.
"""
import os
import sys
import warnings

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import popcorn_b200 as pb  # noqa: E402
from popcorn_b200 import ops, weights  # noqa: E402
from popcorn_b200.model import popcorn as pm  # noqa: E402
from oracle import popcorn_oracle as po  # noqa: E402


def rel(a, b, floor=1e-2):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float(((a - b).abs() / torch.clamp(b.abs(), min=floor * float(b.abs().max()))).max())


def torch_head(params, feats, builtup, idx, B, HW, dtype):
    """autograd reference on the GPU over the SAME gathered features."""
    ps = [p.detach().to(dtype).requires_grad_(True) for p in params]
    C = feats.shape[1]
    flat = feats.permute(1, 0, 2, 3).reshape(C, -1).to(dtype)
    h = flat[:, idx.long()].t()
    for i in range(3):
        h = F.relu(F.linear(h, ps[2 * i].flatten(1), ps[2 * i + 1]))
    o = F.linear(h, ps[6].flatten(1), ps[7])[:, 0]
    scale = F.relu(o)
    dens = scale * builtup.reshape(-1).to(dtype)[idx.long()]
    b = (idx.long() // HW)
    pop = torch.zeros(B, dtype=dtype, device=feats.device).index_add_(0, b, dens)
    return ps, scale, pop


def main():
    torch.cuda.set_device(0)
    sd = po.random_state_dict(seed=1600)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = pb.POPCORN(6, occupancymodel=True, pretrained=False, biasinit=0.9407, sentinelbuildings=True, device="cuda")
    model.load_state_dict(sd)
    model.train()
    cases = [("smoke", 96, 128, [(10, 70, 20, 100), (30, 90, 8, 64)], [2500.0, 9000.0]),
             ("n%128=0", 96, 128, [(0, 64, 0, 64), (0, 32, 0, 128)], [2500.0, 9000.0]),
             ("n%128=36", 96, 128, [(10, 71, 20, 96), (30, 90, 8, 64)], [3500.0, 12000.0])]
    for name, H, W, boxes, ys in cases:
        B = 2
        x = po.synthetic_input(H, W, seed=3, B=B)
        admin = torch.zeros(B, H, W)
        for b, (r0, r1, c0, c1) in enumerate(boxes):
            admin[b, r0:r1, c0:c1] = float(4 + 5 * b)
        cidx = torch.tensor([4, 9])
        y = torch.tensor(ys)
        torch.manual_seed(7)
        grid = po.sparsity_grid(H, W)
        sdg = {k: (v.clone().requires_grad_(True) if k.startswith("head.") else v) for k, v in sd.items()}
        r = po.forward(sdg, {"input": x, "admin_mask": admin, "census_idx": cidx}, padding=False, sparse=True, grid=grid)
        po.train_loss(r, y).backward()
        n_ref = int(r["mask"].sum())
        print(f"=== case {name}: n = {n_ref} (n % 128 = {n_ref % 128}); oracle popcount {r['popcount'].tolist()} y {ys}")
        for tc in (True, False):
            pm.USE_TENSOR_CORE_HEAD = tc
            for p in model.parameters():
                p.grad = None
            torch.manual_seed(7)
            inp = {"input": x.cuda(), "admin_mask": admin.cuda(), "census_idx": cidx.cuda()}
            o = model(inp, train=True, padding=False, encoder_no_grad=True, unet_no_grad=True, sparse=True)
            loss = po.train_loss(o, y.cuda())
            loss.backward()
            idx, n_dev = model._last_compaction
            n = int(n_dev.item())
            idx = idx[:n]
            mask = torch.zeros(B * H * W, dtype=torch.bool)
            mask[idx.long().cpu()] = True
            print(f" [head {'tcgen05' if tc else 'simt'}] n {n} mask==oracle {bool(torch.equal(mask.view(B, H, W), r['mask']))} "
                  f"idx sorted {bool((idx[1:] > idx[:-1]).all())}")
            # ---- A: forward pieces vs the CPU oracle
            bu = inp["building_counts"]
            with torch.no_grad():
                feats = ops.dda_forward(model._dda_pack("unetmodel"), inp["input"], (0, 0, 0, 0), ops.PC_DDA_FEATURES)
                ref_feats = po.unet_features(sd, x, False)
                ref_bu = po.building_score(sd, x)
            print(f"   A feats rel(max-floor) {rel(feats, ref_feats, 1.0):.2e}  builtup {rel(bu, ref_bu, 1.0):.2e}  "
                  f"scale_sel {rel(o['scale'], r['scale'], 1e-3):.2e}  popcount {rel(o['popcount'], r['popcount'], 1.0):.2e}  "
                  f"loss {float(loss):.6f} vs {float(po.train_loss(r, y)):.6f}")
            print(f"     feats absmax {float(feats.abs().max()):.3e} builtup min/max {float(bu.min()):.3e}/{float(bu.max()):.3e} "
                  f"scale>0 frac {float((o['scale'] > 0).float().mean()):.3f} scale max {float(o['scale'].max()):.3e}")
            # ---- C: end-to-end per parameter (what smoke() asserts)
            for k, p in model.named_parameters():
                if k.startswith("head."):
                    print(f"   C {k:14s} rel {rel(p.grad, sdg[k].grad):.3e}  |ref|max {float(sdg[k].grad.abs().max()):.3e}")
            # ---- B: the backward kernel alone vs torch autograd on the GPU over OUR feats
            params = model._head_params()
            hpack = weights.pack_head(dict(zip([f"head.{i}.{t}" for i in (0, 2, 4, 6) for t in ("weight", "bias")], params))).detach()
            for dt in (torch.float32, torch.float64):
                ps, scale_t, pop_t = torch_head(params, feats, bu, idx, B, H * W, dt)
                for term in ("pop", "scale", "both"):
                    for q in ps:
                        q.grad = None
                    lt = 0
                    if term in ("pop", "both"):
                        lt = lt + F.l1_loss(torch.log(pop_t + 1), torch.log(y.cuda().to(dt) + 1))
                    if term in ("scale", "both"):
                        lt = lt + 0.01 * scale_t.abs().mean()
                    (lt * 100.0).backward(retain_graph=True)
                    # the same upstream gradients, handed to the kernel
                    g_pop = torch.zeros(B, device="cuda")
                    g_sel = None
                    if term in ("pop", "both"):
                        pt = pop_t.detach().float()
                        g_pop = 100.0 * 0.5 * torch.sign(torch.log(pt + 1) - torch.log(y.cuda() + 1)) / (pt + 1)
                    if term in ("scale", "both"):
                        g_sel = (100.0 * 0.01 / n) * torch.sign(scale_t.detach().float())
                    gp = ops.head_sparse_backward(hpack, feats, bu, idx, n_dev, n, g_pop, 0.0, g_sel)
                    g = weights.unpack_head_grad(gp, 16)
                    errs = [rel(g[k], q.grad) for k, q in zip([f"head.{i}.{t}" for i in (0, 2, 4, 6) for t in ("weight", "bias")], ps)]
                    print(f"   B kernel vs torch-{str(dt)[6:]:8s} term {term:5s}: max rel {max(errs):.3e}  per-param "
                          + " ".join(f"{e:.1e}" for e in errs))
            # torch GPU autograd (fp64, our feats) vs CPU oracle grads: isolates feature differences
            ps, scale_t, pop_t = torch_head(params, feats, bu, idx, B, H * W, torch.float64)
            lt = (F.l1_loss(torch.log(pop_t + 1), torch.log(y.cuda().double() + 1)) + 0.01 * scale_t.abs().mean()) * 100
            lt.backward()
            errs = [rel(q.grad, sdg[k].grad) for k, q in zip([f"head.{i}.{t}" for i in (0, 2, 4, 6) for t in ("weight", "bias")], ps)]
            print("   D torch-f64(our feats) vs CPU oracle: " + " ".join(f"{e:.1e}" for e in errs))
    pm.USE_TENSOR_CORE_HEAD = True


if __name__ == "__main__":
    main()
