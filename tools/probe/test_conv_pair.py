"""Round-2 bring-up check of tools/probe/conv_pair.cu (fp16-pair operands) and tools/probe/conv_ss.cu (3xTF32, raw fp32 A from
shared memory) — the two stager-free conv candidates — needs a B200.

    bash tools/probe/build_conv_pair.sh && timeout 120 python tools/probe/test_conv_pair.py [pair|ss]

Each case runs in the same process; a trap (barrier protocol mistake) poisons the context, so the first failure ends the run.
Checks, per layer shape of the DDA UNet:
  * planar fp32 output  == relu(conv2d(x_pair, w_pair) + b) computed in fp64 on the SAME fp16-pair operands   (tolerance 2e-5 rel:
    fp32 accumulation order only);
  * pair output         == the pair packing of that result (to 2^-21 relative);
  * against the true fp32 conv the error is the 22-bit operand rounding (~1e-6 relative), what profiles/r1c_precision_study.md
    budgets for.
"""
import ctypes as C
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "popcorn_b200", "libpopcorn_b200_probe.so")


def bf16_pieces(x: torch.Tensor):
    """(h1, h2) fp16 pieces of x (the name predates the switch from bf16 to fp16 pairs: profiles/r1c_precision_study.md)."""
    b1 = x.clamp(-65504.0, 65504.0).to(torch.float16)
    b2 = (x.clamp(-65504.0, 65504.0) - b1.float()).to(torch.float16)
    return b1, b2


def to_pair(x: torch.Tensor) -> torch.Tensor:
    """[C,H,W] fp32 -> [C/4,H,W,4] int32 words b1 | b2 << 16."""
    b1, b2 = bf16_pieces(x)
    w = (b1.view(torch.int16).to(torch.int32) & 0xFFFF) | (b2.view(torch.int16).to(torch.int32) << 16)
    Cc, H, W = x.shape
    return w.view(Cc // 4, 4, H, W).permute(0, 2, 3, 1).contiguous()


def from_pair(w: torch.Tensor) -> torch.Tensor:
    """[C/4,H,W,4] int32 -> [C,H,W] fp32 (b1 + b2)."""
    b1 = (w & 0xFFFF).to(torch.int16).view(torch.float16).float()
    b2 = (w >> 16).to(torch.int16).view(torch.float16).float()
    v = b1 + b2
    q, H, W, _ = w.shape
    return v.permute(0, 3, 1, 2).reshape(4 * q, H, W).contiguous()


def pair_value(x: torch.Tensor) -> torch.Tensor:
    b1, b2 = bf16_pieces(x)
    return b1.double() + b2.double()


def to_c4(x: torch.Tensor) -> torch.Tensor:
    """[C,H,W] fp32 -> [C/4,H,W,4] fp32 chunks (conv_ss.cu's layout)."""
    Cc, H, W = x.shape
    return x.view(Cc // 4, 4, H, W).permute(0, 2, 3, 1).contiguous()


def from_c4(c: torch.Tensor) -> torch.Tensor:
    q, H, W, _ = c.shape
    return c.permute(0, 3, 1, 2).reshape(4 * q, H, W).contiguous()


def run_case_ss(lib, cin_a, cin_b, cout, H, W, pool=False, b_shape=None, b_off=(0, 0), tile_rows=0, seed=0):
    """conv_ss.cu: 3xTF32 — compared with the exact conv (fp64); expected error ~1e-6 relative, bar 1e-5."""
    g = torch.Generator().manual_seed(seed)
    cin = cin_a + cin_b
    xa = torch.randn(cin_a, H, W, generator=g) * 2
    bH, bW = b_shape or (H, W)
    xb = torch.randn(cin_b, bH, bW, generator=g) * 2 if cin_b else None
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    flat = torch.cat([w.permute(1, 2, 3, 0).reshape(-1), bias]).contiguous()
    x_all = xa
    if cin_b:
        full_b = torch.zeros(cin_b, H, W)
        full_b[:, b_off[0]:b_off[0] + bH, b_off[1]:b_off[1] + bW] = xb
        x_all = torch.cat([xa, full_b], 0)
    ref = F.relu(F.conv2d(x_all[None].double(), w.double(), bias.double(), padding=1))[0]
    dev = "cuda"
    a_c = to_c4(xa).to(dev)
    b_c = to_c4(xb).to(dev) if cin_b else None
    out_c4 = torch.full((cout // 4, H, W, 4), float("nan"), device=dev)
    out_planar = torch.full((cout, H, W), float("nan"), device=dev)
    pool_c4 = torch.full((cout // 4, H // 2, W // 2, 4), float("nan"), device=dev) if pool else None
    rc = lib.pc_probe_conv3x3_ss(a_c.data_ptr(), cin_a // 4, H, W, 0, 0, b_c.data_ptr() if cin_b else None, cin_b // 4, bH, bW,
                                 b_off[0], b_off[1], flat.data_ptr(), cout, 1, H, W, out_c4.data_ptr(), out_planar.data_ptr(),
                                 pool_c4.data_ptr() if pool else None, tile_rows, None)
    assert rc == 0, f"pc_probe_conv3x3_ss returned {rc}"
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    e_planar = float((out_planar.cpu().double() - ref).abs().max()) / scale
    e_c4 = float((from_c4(out_c4.cpu()).double() - ref).abs().max()) / scale
    same = torch.equal(from_c4(out_c4.cpu()), out_planar.cpu())
    msg = f"ss cin {cin_a}+{cin_b} cout {cout} {H}x{W} TR {tile_rows}: planar {e_planar:.2e}  chunks {e_c4:.2e}  planar == chunks {same}"
    ok = e_planar < 1e-5 and same
    if pool:
        pref = F.max_pool2d(ref[None], 2)[0]
        e_pool = float((from_c4(pool_c4.cpu()).double() - pref).abs().max()) / scale
        msg += f"  pool {e_pool:.2e}"
        ok = ok and e_pool < 1e-5
    print(("OK   " if ok else "FAIL ") + msg, flush=True)
    return ok


def run_c4_helpers(lib):
    """tools/probe/c4_kernels.cu: layout converters (exact) and ConvTranspose2d k2 s2 on chunks (vs F.conv_transpose2d)."""
    vp, i, ll = C.c_void_p, C.c_int, C.c_longlong
    lib.pc_probe_planar_to_c4.restype = i
    lib.pc_probe_planar_to_c4.argtypes = [vp, ll, i, i, i, i, vp, vp]
    lib.pc_probe_c4_to_planar.restype = i
    lib.pc_probe_c4_to_planar.argtypes = [vp, i, i, i, vp, ll, i, vp]
    lib.pc_probe_convt2x2_c4.restype = i
    lib.pc_probe_convt2x2_c4.argtypes = [vp, vp, i, i, i, vp, vp]
    ok = True
    for Cc, H, W in ((8, 37, 301), (16, 64, 128)):
        x = torch.randn(Cc, H, W).cuda()
        c4 = torch.empty(Cc // 4, H, W, 4, device="cuda")
        back = torch.empty_like(x)
        assert lib.pc_probe_planar_to_c4(x.data_ptr(), H * W, W, Cc, H, W, c4.data_ptr(), None) == 0
        assert lib.pc_probe_c4_to_planar(c4.data_ptr(), Cc, H, W, back.data_ptr(), H * W, W, None) == 0
        torch.cuda.synchronize()
        good = torch.equal(c4.cpu(), to_c4(x.cpu())) and torch.equal(back, x)
        print(("OK   " if good else "FAIL ") + f"planar <-> c4 {Cc}x{H}x{W}", flush=True)
        ok &= good
        wt = torch.randn(Cc, Cc, 2, 2) * 0.3
        bt = torch.randn(Cc) * 0.1
        pack = torch.cat([wt.permute(0, 2, 3, 1).reshape(-1), bt]).contiguous().cuda()       # [ci][dy][dx][co] + bias
        out = torch.empty(Cc // 4, 2 * H, 2 * W, 4, device="cuda")
        assert lib.pc_probe_convt2x2_c4(c4.data_ptr(), pack.data_ptr(), Cc, H, W, out.data_ptr(), None) == 0
        torch.cuda.synchronize()
        ref = F.conv_transpose2d(x.cpu().double()[None], wt.double(), bt.double(), stride=2)[0]
        e = float((from_c4(out.cpu()).double() - ref).abs().max() / ref.abs().max())
        print(("OK   " if e < 1e-6 else "FAIL ") + f"convt2x2_c4<{Cc}> {H}x{W}: {e:.2e}", flush=True)
        ok &= e < 1e-6
    return ok


def run_case(lib, cin_a, cin_b, cout, H, W, pool=False, b_shape=None, b_off=(0, 0), tile_rows=0, seed=0):
    g = torch.Generator().manual_seed(seed)
    cin = cin_a + cin_b
    xa = torch.randn(cin_a, H, W, generator=g) * 2
    bH, bW = b_shape or (H, W)
    xb = torch.randn(cin_b, bH, bW, generator=g) * 2 if cin_b else None
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    flat = torch.cat([w.permute(1, 2, 3, 0).reshape(-1), bias]).contiguous()        # [cin][ky][kx][cout] + bias
    # reference on the same pair operands (fp64), and the plain fp32 conv
    full_b = None
    if cin_b:
        full_b = torch.zeros(cin_b, H, W)
        oy, ox = b_off
        full_b[:, oy:oy + bH, ox:ox + bW] = xb
    x_all = xa if not cin_b else torch.cat([xa, full_b], 0)
    ref = F.relu(F.conv2d(pair_value(x_all)[None], pair_value(w), bias.double(), padding=1))[0]
    ref32 = F.relu(F.conv2d(x_all[None].double(), w.double(), bias.double(), padding=1))[0]
    dev = "cuda"
    a_p = to_pair(xa).to(dev)
    b_p = to_pair(xb).to(dev) if cin_b else None
    out_pair = torch.full((cout // 4, H, W, 4), -1, dtype=torch.int32, device=dev)
    out_planar = torch.full((cout, H, W), float("nan"), device=dev)
    pool_pair = torch.full((cout // 4, H // 2, W // 2, 4), -1, dtype=torch.int32, device=dev) if pool else None
    rc = lib.pc_probe_conv3x3_pair(a_p.data_ptr(), cin_a // 4, H, W, 0, 0, b_p.data_ptr() if cin_b else None, cin_b // 4, bH, bW,
                                   b_off[0], b_off[1], flat.data_ptr(), cout, 1, H, W, out_pair.data_ptr(), out_planar.data_ptr(),
                                   pool_pair.data_ptr() if pool else None, tile_rows, None)
    assert rc == 0, f"pc_probe_conv3x3_pair returned {rc} ({torch.cuda.get_device_name(0)})"
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    e_planar = float((out_planar.cpu().double() - ref).abs().max()) / scale
    e_pair = float((from_pair(out_pair.cpu()).double() - ref).abs().max()) / scale
    e_fp32 = float((out_planar.cpu().double() - ref32).abs().max()) / scale
    msg = f"cin {cin_a}+{cin_b} cout {cout} {H}x{W} TR {tile_rows}: planar {e_planar:.2e}  pair {e_pair:.2e}  vs fp32 conv {e_fp32:.2e}"
    ok = e_planar < 2e-5 and e_pair < 2e-5
    if pool:
        pref = F.max_pool2d(ref[None], 2)[0]
        e_pool = float((from_pair(pool_pair.cpu()).double() - pref).abs().max()) / scale
        msg += f"  pool {e_pool:.2e}"
        ok = ok and e_pool < 2e-5
    print(("OK   " if ok else "FAIL ") + msg, flush=True)
    return ok


if __name__ == "__main__":
    assert torch.cuda.is_available(), "needs a CUDA device"
    lib = C.CDLL(LIB)
    vp, i = C.c_void_p, C.c_int
    lib.pc_probe_conv3x3_pair.restype = i
    lib.pc_probe_conv3x3_pair.argtypes = [vp, i, i, i, i, i, vp, i, i, i, i, i, vp, i, i, i, i, vp, vp, vp, i, vp]
    lib.pc_probe_conv3x3_ss.restype = i
    lib.pc_probe_conv3x3_ss.argtypes = lib.pc_probe_conv3x3_pair.argtypes
    ok = True
    if len(sys.argv) > 1 and sys.argv[1] == "ss":
        run_case = run_case_ss                                                # same shapes, the 3xTF32 SS-form kernel
        ok &= run_c4_helpers(lib)
    ok &= run_case(lib, 8, 0, 8, 64, 128)                                   # one tile, no halo columns outside
    ok &= run_case(lib, 8, 0, 8, 70, 200, tile_rows=32)                      # several tiles, ragged right edge, ring wraps
    ok &= run_case(lib, 8, 0, 8, 37, 130, tile_rows=32)                      # odd height
    ok &= run_case(lib, 8, 0, 16, 96, 256)
    ok &= run_case(lib, 16, 0, 16, 96, 256)
    ok &= run_case(lib, 8, 0, 8, 128, 256, pool=True)
    ok &= run_case(lib, 16, 0, 16, 64, 128, pool=True)
    ok &= run_case(lib, 8, 8, 8, 96, 160)                                    # skip + upsampled branch
    ok &= run_case(lib, 16, 16, 8, 67, 131, b_shape=(66, 130), b_off=(0, 0))  # F.pad of the Up block: branch smaller, zero outside
    ok &= run_case(lib, 8, 0, 8, 512, 2048)                                   # many tiles per CTA
    print("ALL OK" if ok else "SOME FAILED")
    sys.exit(0 if ok else 1)
