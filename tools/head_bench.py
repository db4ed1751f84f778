"""Dense tcgen05 head at a bench-like size, timed with CUDA events.  Development tool."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from popcorn_b200 import ops, weights
from oracle import popcorn_oracle as po

H, W = 3840, 8192
sd = po.random_state_dict()
hp = weights.pack_head_tc(sd).cuda()
feats = torch.randn(1, 16, H, W, device="cuda")
bu = torch.rand(1, 1, H, W, device="cuda")
for _ in range(2):
    ops.head_dense_forward(hp, feats, bu, None, None, None, tc=True)
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    ops.head_dense_forward(hp, feats, bu, None, None, None, tc=True)
e.record()
torch.cuda.synchronize()
ms = a.elapsed_time(e) / 5
px = H * W
print(f"head_tc dense {H}x{W}: {ms:.3f} ms  {px / ms / 1e6:.1f} Gpx/s  {18688 * px / ms / 1e9:.1f} TFLOP/s (algorithmic)  "
      f"clk/tile/SM {ms * 1e-3 * 148 * 1.9e9 / (px / 128):.0f} (pipe 1728)")
