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

from popcorn_b200 import _lib
L = _lib.lib()
if hasattr(L, "pc_debug_head_counters"):      # probe build (-DPC_HEAD_PROBE=1): cycle counters of CTA 0
    import ctypes
    buf = (ctypes.c_longlong * 32)()
    ops.head_dense_forward(hp, feats, bu, None, None, None, tc=True)
    L.pc_debug_head_counters(buf)
    c = list(buf)
    n = max(c[7], 1)
    print("  worker (ctx 0, warp 0), per tile: stage+handover, wait_d1, epi1+handover, wait_d2, epi2+handover, wait_d3, output layer+stores =",
          [round(v / n) for v in c[0:7]], " tiles", c[7])
    print("  detail: epi1 math, epi1 handover, locate+feature loads | output: ld+dot, bar1, combine+stores, bar2 =", [round(v / n) for v in c[8:15]])
    m = max(c[23], 1)
    print("  issuer, per issue: L1 wait / issue, L2-3 wait / issue (cycles summed over both contexts, per tile-layer) =",
          [round(c[16] / (m / 3)), round(c[17] / (m / 3)), round(c[18] / (2 * m / 3)), round(c[19] / (2 * m / 3))], " issues", c[23])
