"""Timeline of the warp-specialised tensor-core conv pipeline (probe build, -DPC_TC_PROBE=1): clock64 stamps of CTA (0,0)
for batches 64..127.  Development tool:  POPCORN_B200_LIB=.../libpopcorn_b200_probe.so python tools/conv_tc_trace.py [cin cout]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from popcorn_b200 import _lib

L = _lib.lib()
cin, cout = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) >= 3 else (8, 8)
H, W = 2048, 4096
st = torch.cuda.current_stream().cuda_stream
x = torch.randn(cin, H, W, device="cuda")
w = torch.randn(cout, cin, 3, 3) * 0.2
b = torch.randn(cout)
flat = torch.cat([w.permute(1, 2, 3, 0).reshape(-1), b]).contiguous()
img = torch.zeros(L.pc_conv_tc_layer_floats(cin, cout))
_lib.check(L.pc_conv_tc_pack_layer(flat.data_ptr(), cin, cout, img.data_ptr()))
flat_d, img_d = flat.cuda(), img.cuda()
out = torch.empty(cout, H, W, device="cuda")
for _ in range(3):
    _lib.check(L.pc_test_conv3x3(x.data_ptr(), cin, H, W, 0, 0, 0, None, 0, 0, 0, 0, 0, flat_d.data_ptr(), cout, H, W,
                                 out.data_ptr(), None, img_d.data_ptr(), st))
buf = (ctypes.c_longlong * 4096)()
L.pc_debug_tc_counters(buf, 2)
t = list(buf)
ev = lambda role, batch, e: t[(role * 64 + batch) * 8 + e]
base = min(v for v in t if v > 0)
print("batch | stager0: start s_full empty_a d_empty->stage done | stager1 ... | issuer0: start ready issued committed | issuer1 | epi0/1: start ready done")
for B in range(0, 40):
    row = [f"{B + 64:4d}"]
    for role, n in ((0, 5), (1, 5), (2, 4), (3, 4), (4, 3), (5, 3)):
        row.append(" ".join(f"{(ev(role, B, e) - base) if ev(role, B, e) else -1:7d}" for e in range(n)))
    print(" | ".join(row))
