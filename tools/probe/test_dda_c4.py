"""Round-2 A/B: one DDA feature pass in the chunk layout (tools/probe/dda_c4.cu on conv_ss.cu) against the shipped
pc_dda_forward — same weights, same input; compares the 16 feature planes and times both.  Needs a B200.

    bash tools/probe/build_conv_pair.sh && timeout 300 python tools/probe/test_dda_c4.py [H W]
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from popcorn_b200 import ops, synthetic, weights  # noqa: E402

if __name__ == "__main__":
    assert torch.cuda.is_available(), "needs a CUDA device"
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 2048)
    probe = C.CDLL(os.path.join(ROOT, "popcorn_b200", "libpopcorn_b200_probe.so"))
    probe.pc_probe_dda_features_c4.restype = C.c_int
    probe.pc_probe_dda_features_c4.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_void_p]
    sd = synthetic.random_state_dict(seed=1600)
    pack = weights.pack_dda(sd, "unetmodel")                      # fp32 section first, tcgen05 images after it
    pack_host = pack.cpu().contiguous()
    x = torch.randn(1, 6, H, W, device="cuda")
    iters = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pack_dev = pack.cuda()
    with torch.no_grad():
        ref = ops.dda_forward(pack_dev, x, (0, 0, 0, 0), ops.PC_DDA_FEATURES).clone()
        for _ in range(2):
            ops.dda_forward(pack_dev, x, (0, 0, 0, 0), ops.PC_DDA_FEATURES)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            ops.dda_forward(pack_dev, x, (0, 0, 0, 0), ops.PC_DDA_FEATURES)
        e1.record()
        torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1) / iters
    out = torch.full((16, H, W), float("nan"), device="cuda")
    ms = C.c_float(0)
    rc = probe.pc_probe_dda_features_c4(pack_host.data_ptr(), x.data_ptr(), H, W, out.data_ptr(), iters, C.byref(ms), None)
    assert rc == 0, f"pc_probe_dda_features_c4 returned {rc}"
    torch.cuda.synchronize()
    d = (out - ref[0]).abs()
    rel = float(d.max() / ref.abs().max())
    print(f"{H}x{W}: shipped pc_dda_forward {ms_ref:.3f} ms, chunk-layout pass {ms.value:.3f} ms ({ms_ref / max(ms.value, 1e-9):.2f}x); "
          f"features max |diff| / max|ref| = {rel:.2e}  ({'OK' if rel < 1e-4 else 'MISMATCH'})")
    sys.exit(0 if rel < 1e-4 else 1)
