"""One launch each of the hot kernels at bench-like sizes (for `ncu --set full`).  Development tool."""
import os, sys, warnings
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from popcorn_b200 import ops, weights
from oracle import popcorn_oracle as po
H, W = 2048, 4096
sd = po.random_state_dict()
x = torch.randn(1, 6, H, W, device="cuda")
pack = weights.pack_dda(sd, "unetmodel").cuda()
hp = weights.pack_head(sd).cuda()
for _ in range(2):
    feats = ops.dda_forward(pack, x, (0, 0, 0, 0), 0)
bu = torch.rand(1, 1, H, W, device="cuda")
hp_tc = weights.pack_head_tc(sd).cuda()
for _ in range(2):
    dens, scale = ops.head_dense_forward(hp_tc, feats, bu, None, None, None, tc=True)
ids = po.synthetic_regions(H, W, 400).cuda().contiguous()
n = 1 << 27
d = torch.rand(n, device="cuda"); big = ids.reshape(-1).repeat(n // ids.numel() + 1)[:n].contiguous()
for _ in range(2):
    ops.region_sum(d, big, 401)
torch.cuda.synchronize()
