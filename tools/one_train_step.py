"""One census train step (sparse head forward + backward) plus one dense eval forward with census sums and a tile accumulation, on a
small or a config-3-sized batch: the target of the compute-sanitizer and `ncu --set full` runs (development tool).
    python tools/one_train_step.py [small|config3] [finetune]"""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import popcorn_b200 as pb  # noqa: E402
from popcorn_b200 import ops  # noqa: E402
from popcorn_b200 import synthetic as sy  # noqa: E402

size = sys.argv[1] if len(sys.argv) > 1 else "small"
ft = len(sys.argv) > 2 and sys.argv[2] == "finetune"
dev = torch.device("cuda", 0)
sd, _ = bench.bench_weights()
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    model = pb.POPCORN(6, occupancymodel=True, pretrained=False, biasinit=0.9407, sentinelbuildings=True, device=dev)
model.load_state_dict(sd)
if size == "config3":
    x, admin, cidx, y = bench._train_batch(dev)
else:
    B, H, W = 2, 96, 160
    x = bench.synth_raster_slab(B * H, W, 5, dev).view(6, B, H, W).permute(1, 0, 2, 3).contiguous()
    admin = torch.zeros(B, H, W, device=dev)
    admin[0, 10:70, 20:100] = 4.0
    admin[1, 30:90, 8:150] = 9.0
    cidx = torch.tensor([4, 9], device=dev)
    y = torch.tensor([2500.0, 9000.0], device=dev)
model.train()
for it in range(2):
    out = model({"input": x, "admin_mask": admin, "census_idx": cidx}, train=True, padding=False, encoder_no_grad=not ft,
                unet_no_grad=not ft, sparse=True)
    sy.census_loss(out, y).backward()
model.eval()
with torch.no_grad():
    o = model({"input": x[:1]}, padding=False)
    ids = (torch.arange(x.shape[2] * x.shape[3], device=dev) % 7).to(torch.int32).view(x.shape[2], x.shape[3])
    s = ops.region_sum(o["popdensemap"][0].contiguous(), ids, 7)
    Hh, Ww = x.shape[2], x.shape[3]
    maps = [torch.zeros(Hh, Ww, device=dev) for _ in range(4)] + [torch.zeros(Hh, Ww, dtype=torch.int16, device=dev)]
    ops.accumulate_tile(o["popdensemap"][0], o["scale"][0], (8, Hh - 8), (8, Ww - 8), maps, 0, 0)
    ops.accumulate_tile(o["popdensemap"][0], o["scale"][0], (8, Hh - 8), (8, Ww - 8), maps, 0, 0)
    ops.finalize_map(maps)
torch.cuda.synchronize()
print("ok", float(s.sum()), float(out["popcount"].sum()), float(maps[0].sum()))
