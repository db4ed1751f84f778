"""How fast is pinned H2D / D2H on this box, and how does the streamed engine overlap it?  Development tool."""
import os, sys, time, warnings
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import popcorn_b200 as pb
from popcorn_b200 import ops, country as ct
from oracle import popcorn_oracle as po
H, W = 15104, 17216
dev = torch.device("cuda")
host = torch.empty(6, H, W, dtype=torch.float32, pin_memory=True)
host.normal_()
d = torch.empty(6, 3840, W, device=dev)
torch.cuda.synchronize()
for rows in (2048, 3840):
    t0 = time.perf_counter()
    for r0 in range(0, H - rows, rows):
        ops.copy_window_h2d(d[:, :rows], host[:, r0:r0 + rows])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n = len(range(0, H - rows, rows)) * 6 * rows * W * 4
    print(f"H2D memcpy2d rows={rows}: {n/dt/1e9:.1f} GB/s")
t0 = time.perf_counter()
for r0 in range(0, H - 3840, 3840):
    d.copy_(host[:, r0:r0 + 3840], non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"H2D torch copy_ contiguous planes: {3 * 6 * 3840 * W * 4/dt/1e9:.1f} GB/s")
m = torch.empty(H, W, device=dev); hm = torch.empty(H, W, dtype=torch.float32, pin_memory=True)
t0 = time.perf_counter(); ops.copy_d2h(hm, m); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"D2H map: {H*W*4/dt/1e9:.1f} GB/s")
sd = po.random_state_dict()
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    model = pb.POPCORN(6, occupancymodel=True, sentinelbuildings=True, device=dev)
model.load_state_dict(sd); model.eval()
ids = torch.zeros(H - 256, W, dtype=torch.int32, device=dev)
for rps in (1, 2):
    eng = ct.CountryEngine([model], H, W, rows_per_strip=rps)
    with torch.no_grad():
        for _ in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            o = eng.run(host, ids, 401)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            ops.copy_d2h(hm[: o["map"].shape[0]], o["map"]); s = o["sums"].cpu()
            torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"rps={rps}: streamed run {1e3*(t1-t0):.0f} ms, + d2h {1e3*(t2-t1):.0f} ms, windows={len(eng.windows)}, h2d={eng.h2d_bytes/1e9:.2f} GB")
