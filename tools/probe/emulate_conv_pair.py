"""CPU emulation of tools/probe/conv_pair.cu's and conv_ss.cu's data flow — index math only, runs anywhere:  python tools/probe/emulate_conv_pair.py

Re-implements in numpy what the kernel's three roles do — the host weight packing ([w1 | w2][kx][chunk][48 rows][4 ch x (w, w)]),
the staged 136-pixel rows in 16-byte chunks, the UMMA windows (A = 128 pixels x K16 at a +kx pixel shift, B = 16*n rows starting
at the ky block of the first output row of the run), the 8-slot accumulator ring with its wrap, halo rows and zero padding — and
checks the result against conv2d on the same fp16-pair operands.  What it cannot check is the hardware's reading of the
descriptors (LBO / SBO / element order inside a chunk): that is tools/probe/pair_probe.cu's job on a B200."""
import sys, numpy as np, torch, torch.nn.functional as F
import os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_conv_pair import to_pair, from_pair, pair_value, bf16_pieces
PBOX, PBROWS, PND = 136, 48, 8

def bf(x):  # fp16 RN as float32
    return torch.tensor(x).to(torch.float16).float().numpy()

def pack_layer(flat, cin, cout):
    cq = cin // 4; half = 3 * cq * PBROWS * 8
    img = np.zeros(2 * half, dtype=np.float32)   # hold bf16 values as floats
    for kx in range(3):
        for q in range(cq):
            for ky in range(3):
                for co in range(cout):
                    for e in range(4):
                        w = flat[(((q * 4 + e) * 3 + ky) * 3 + kx) * cout + co]
                        w1 = bf(np.float32(w)); w2 = bf(np.float32(w) - w1)
                        at = ((kx * cq + q) * PBROWS + (2 - ky) * 16 + co) * 8 + 2 * e
                        img[at] = img[at + 1] = w1
                        img[half + at] = img[half + at + 1] = w2
    return img, half

def emulate(xa, w, bias, H, W, TR):
    cin, cout = xa.shape[0], w.shape[0]; cq = cin // 4
    flat = torch.cat([w.permute(1, 2, 3, 0).reshape(-1), bias]).numpy()
    img, half = pack_layer(flat, cin, cout)
    b1, b2 = bf16_pieces(xa)
    # pair tensor as floats [cq][H][W][8] = 4 ch x (b1,b2)
    P = torch.stack([b1.float(), b2.float()], -1).view(cq, 4, H, W, 2).permute(0, 2, 3, 1, 4).reshape(cq, H, W, 8).numpy()
    out = np.zeros((cout, H, W), dtype=np.float64)
    tiles_x = (W + 127) // 128; tiles_y = (H + TR - 1) // TR
    ring = np.zeros((PND, 128, 16)); g0 = 0
    for tile in range(tiles_x * tiles_y):
        ty, tx = divmod(tile, tiles_x); x0, y0 = tx * 128, ty * TR
        nrows = (min(H - y0, TR) + 1) & ~1
        for r in range(-1, nrows + 1):
            stage = np.zeros((cq, PBOX, 8))
            y = y0 + r
            for p in range(PBOX):
                x = x0 - 1 + p
                if 0 <= x < W and 0 <= y < H: stage[:, p] = P[:, y, x]
            lo, hi = max(r - 1, 0), min(r + 1, nrows - 1)
            o = lo
            while o <= hi:
                slot = (g0 + o) % PND; n = min(hi - o + 1, PND - slot); brow = 16 * (o - r + 1)
                for kx in range(3):
                    for j in range(cq // 2):
                        A = np.concatenate([stage[2 * j, kx:kx + 128], stage[2 * j + 1, kx:kx + 128]], 1)       # [128, 16]
                        for im in range(2):
                            rows = []
                            for ch in (2 * j, 2 * j + 1):
                                base = im * half + ((kx * cq + ch) * PBROWS + brow) * 8
                                rows.append(img[base: base + 16 * n * 8].reshape(16 * n, 8))
                            B = np.concatenate(rows, 1)                                                               # [16n, 16]
                            D = A.astype(np.float64) @ B.astype(np.float64).T                                          # [128, 16n]
                            for t in range(n): ring[slot + t] += D[:, 16 * t:16 * t + 16]
                o += n
            if r >= 1:
                orow = r - 1; slot = (g0 + orow) % PND; oy = y0 + orow
                if oy < H:
                    for px in range(128):
                        if x0 + px < W: out[:, oy, x0 + px] = ring[slot, px, :cout] + flat[cin * 9 * cout: cin * 9 * cout + cout]
                ring[slot] = 0
        g0 += nrows
    return np.maximum(out, 0)

torch.manual_seed(0)
for (cin, cout, H, W, TR) in ((8, 8, 37, 130, 32), (16, 16, 20, 140, 8), (8, 16, 10, 64, 4)):
    xa = torch.randn(cin, H, W) * 2; w = torch.randn(cout, cin, 3, 3) * 0.2; bias = torch.randn(cout) * 0.1
    ref = F.relu(F.conv2d(pair_value(xa)[None], pair_value(w), bias.double(), padding=1))[0].numpy()
    got = emulate(xa, w, bias, H, W, TR)
    print(cin, cout, H, W, TR, "max err", np.abs(got - ref).max())
assert torch.equal(from_pair(to_pair(xa)).double(), pair_value(xa))
print("converters ok")


# ---------------------------------------------------------------------------------------------------
# conv_ss.cu (3xTF32, raw fp32 A from shared memory + lo buffer): same geometry, fp32 chunks, K = 8, three products
# ---------------------------------------------------------------------------------------------------
def trunc_tf32(a):
    return (np.asarray(a, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def pack_layer_ss(flat, cin, cout):
    cq = cin // 4; half = 3 * cq * PBROWS * 4
    img = np.zeros(2 * half, dtype=np.float32)
    for kx in range(3):
        for q in range(cq):
            for ky in range(3):
                for co in range(cout):
                    for e in range(4):
                        w = np.float32(flat[(((q * 4 + e) * 3 + ky) * 3 + kx) * cout + co])
                        hi = trunc_tf32(w)
                        at = ((kx * cq + q) * PBROWS + (2 - ky) * 16 + co) * 4 + e
                        img[at] = hi; img[half + at] = w - hi
    return img, half


def emulate_ss(xa, w, bias, H, W, TR):
    cin, cout = xa.shape[0], w.shape[0]; cq = cin // 4
    flat = torch.cat([w.permute(1, 2, 3, 0).reshape(-1), bias]).numpy()
    img, half = pack_layer_ss(flat, cin, cout)
    P = xa.view(cq, 4, H, W).permute(0, 2, 3, 1).contiguous().numpy()          # [cq][H][W][4] fp32 chunks
    out = np.zeros((cout, H, W), dtype=np.float64)
    tiles_x = (W + 127) // 128; tiles_y = (H + TR - 1) // TR
    ring = np.zeros((PND, 128, 16)); g0 = 0
    for tile in range(tiles_x * tiles_y):
        ty, tx = divmod(tile, tiles_x); x0, y0 = tx * 128, ty * TR
        nrows = (min(H - y0, TR) + 1) & ~1
        for r in range(-1, nrows + 1):
            raw = np.zeros((cq, PBOX, 4), dtype=np.float32)
            y = y0 + r
            for p in range(PBOX):
                x = x0 - 1 + p
                if 0 <= x < W and 0 <= y < H: raw[:, p] = P[:, y, x]
            a_hi = trunc_tf32(raw)                              # what the tensor core reads from the raw row
            a_lo = trunc_tf32(raw - a_hi)                       # the lo buffer, as the tensor core reads it
            lo_, hi_ = max(r - 1, 0), min(r + 1, nrows - 1)
            o = lo_
            while o <= hi_:
                slot = (g0 + o) % PND; n = min(hi_ - o + 1, PND - slot); brow = 16 * (o - r + 1)
                for kx in range(3):
                    for j in range(cq // 2):
                        def A(buf): return np.concatenate([buf[2 * j, kx:kx + 128], buf[2 * j + 1, kx:kx + 128]], 1).astype(np.float64)
                        def B(im):
                            rows = [img[im * half + ((kx * cq + ch) * PBROWS + brow) * 4: im * half + ((kx * cq + ch) * PBROWS + brow + 16 * n) * 4].reshape(16 * n, 4)
                                    for ch in (2 * j, 2 * j + 1)]
                            return trunc_tf32(np.concatenate(rows, 1)).astype(np.float64)
                        D = A(a_hi) @ B(0).T + A(a_lo) @ B(0).T + A(a_hi) @ B(1).T
                        for t in range(n): ring[slot + t] += D[:, 16 * t:16 * t + 16]
                o += n
            if r >= 1:
                orow = r - 1; slot = (g0 + orow) % PND; oy = y0 + orow
                if oy < H:
                    for px in range(128):
                        if x0 + px < W: out[:, oy, x0 + px] = ring[slot, px, :cout] + flat[cin * 9 * cout: cin * 9 * cout + cout]
                ring[slot] = 0
        g0 += nrows
    return np.maximum(out, 0)


for (cin, cout, H, W, TR) in ((8, 8, 37, 130, 32), (16, 16, 20, 140, 8)):
    xa = torch.randn(cin, H, W) * 2; w = torch.randn(cout, cin, 3, 3) * 0.2; bias = torch.randn(cout) * 0.1
    t = lambda v: torch.from_numpy(trunc_tf32(v.numpy()).copy())
    xh, wh = t(xa), t(w)
    xl, wl = t(xa - xh), t(w - wh)
    c = lambda a, b: F.conv2d(a.double()[None], b.double(), None, padding=1)[0]
    ref = F.relu(c(xh, wh) + c(xl, wh) + c(xh, wl) + bias.double().view(-1, 1, 1)).numpy()
    exact = F.relu(c(xa, w) + bias.double().view(-1, 1, 1)).numpy()
    got = emulate_ss(xa, w, bias, H, W, TR)
    print("ss  ", cin, cout, H, W, TR, "max err vs 3xTF32 formula", np.abs(got - ref).max(), " vs exact conv (relative)", np.abs(got - exact).max() / np.abs(exact).max())
