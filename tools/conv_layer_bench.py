"""One 3x3 conv layer (pc_test_conv3x3) at a bench-like size, SIMT vs tcgen05, timed with CUDA events.
Development tool:  python tools/conv_layer_bench.py [cin cout [H W]]   (KB_ONLY=tc|simt, KB_ITERS=n)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from popcorn_b200 import _lib

L = _lib.lib()
cases = [(2, 8), (4, 8), (8, 8), (8, 16), (16, 16)]
if len(sys.argv) >= 3:
    cases = [(int(sys.argv[1]), int(sys.argv[2]))]
H, W = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) >= 5 else (2048, 4096)
only = os.environ.get("KB_ONLY", "")
iters = int(os.environ.get("KB_ITERS", 5))
st = torch.cuda.current_stream().cuda_stream
for cin, cout in cases:
    x = torch.randn(cin, H, W, device="cuda")
    w = torch.randn(cout, cin, 3, 3) * 0.2
    b = torch.randn(cout)
    flat = torch.cat([w.permute(1, 2, 3, 0).reshape(-1), b]).contiguous()
    img = torch.zeros(L.pc_conv_tc_layer_floats(cin, cout))
    _lib.check(L.pc_conv_tc_pack_layer(flat.data_ptr(), cin, cout, img.data_ptr()))
    flat_d, img_d = flat.cuda(), img.cuda()
    out = torch.empty(cout, H, W, device="cuda")
    pool = torch.empty(cout, H // 2, W // 2, device="cuda") if os.environ.get("KB_POOL") else None      # EPI_POOL variant of the layer
    res = {}
    for name, wtc in (("simt", None), ("tc", img_d)):
        if only and only != name:
            continue
        def run():
            _lib.check(L.pc_test_conv3x3(x.data_ptr(), cin, H, W, 0, 0, 0, None, 0, 0, 0, 0, 0, flat_d.data_ptr(), cout, H, W,
                                         out.data_ptr(), None if pool is None else pool.data_ptr(), None if wtc is None else wtc.data_ptr(), st))
        for _ in range(2):
            run()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            run()
        e.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(e) / iters
        res[name] = ms
        if name == "tc" and hasattr(L, "pc_debug_tc_counters"):      # probe build (-DPC_TC_PROBE=1): per-role cycle counters of CTA (0,0)
            import ctypes
            buf = (ctypes.c_longlong * 32)()
            L.pc_debug_tc_counters(buf, 1)
            run()
            L.pc_debug_tc_counters(buf, 1)
            c = list(buf)
            def per(a, n):
                return [round(x / max(n, 1)) for x in a]
            print("   mma issuer0 batches", c[4], " wait_full_a, issue, commit =", per(c[0:3], c[4]))
            print("   stager w0   rows", c[12], " wait_s_full, wait_d_empty, stage, wait_st+arrive, wait_empty_a =", per(c[8:12] + c[13:14], c[12]))
            print("   epilogue w8 rows", c[18], " wait_d_full, ld+zero+arrive =", per(c[16:18], c[18]))
        px = H * W
        print(f"cin {cin:2d} cout {cout:2d} {name:4s}: {ms:7.3f} ms  {px / ms / 1e6:7.1f} Gpx/s  {(cin + cout) * 4 * px / ms / 1e6:7.0f} GB/s  "
              f"{2 * 9 * cin * cout * px / ms / 1e9:6.1f} TFLOP/s  clk/row-tile/SM {ms * 1e-3 * 148 * 1.965e9 / (px / 128):7.0f}", flush=True)
