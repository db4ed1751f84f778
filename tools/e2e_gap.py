"""Where does the end-to-end step lose time against the device-resident step?  (development tool, one GPU)

Times CountryEngine.run on the bench raster in variants that add one host-facing piece at a time:
  resident      fp32 normalised raster in HBM, maps stay on the device                    (= bench.py `value`)
  +download     ... map rows stream to a pinned host map
  raw_device    RawRaster (uint16 S2 + fp32 S1) already in HBM: + the ingest kernels
  raw_host      RawRaster in pinned host memory, prefetch pipeline, no download
  e2e           raw_host + download                                                        (= bench.py `e2e`)
"""
import os
import sys
import time
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import popcorn_b200 as pb  # noqa: E402
from popcorn_b200 import country as ct, ops  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
H, W, name = bench.workload(1)
sd, _ = bench.bench_weights()
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    model = pb.POPCORN(6, occupancymodel=True, pretrained=False, biasinit=0.9407, sentinelbuildings=True, device=dev)
model.load_state_dict(sd)
model.eval()
eng = ct.CountryEngine([model], H, W, merge=True, rows_per_strip=3, rank=0, world=1, first_strip_rows=1, last_strip_rows=1, balance=True,
                       upload_once=True)
i0, i1 = eng.in_rows
lo, hi = eng.out_rows
raster = bench.synth_raster_slab(i1 - i0, W, i0, dev)
ids = bench.synth_ids_slab(H, W, bench.R_REGIONS, lo, hi, dev)
R = bench.R_REGIONS + 1
st2, st1 = ops.DATASET_STATS["sen2springNIR"], ops.DATASET_STATS["sen1"]
rows = raster.shape[1]
host_s2 = torch.empty(4, rows, W, dtype=torch.uint16, pin_memory=True)
host_s1 = torch.empty(2, rows, W, dtype=torch.float32, pin_memory=True)
for dst_plane, c in enumerate((2, 1, 0, 3)):
    v = (raster[c] * st2["std"][c] + st2["mean"][c]).clamp_(0, 10000).round_()
    host_s2[dst_plane].copy_(v.to(torch.int32).to(torch.uint16))
    del v
for c in range(2):
    host_s1[c].copy_(raster[4 + c] * st1["std"][c] + st1["mean"][c])
host_raw = ct.RawRaster(host_s2, host_s1, ops.S2_FILE_TO_RGBN)
dev_raw = ct.RawRaster(host_s2.to(dev), host_s1.to(dev), ops.S2_FILE_TO_RGBN)
host_map = torch.empty(hi - lo, W, dtype=torch.float32, pin_memory=True)
K = int(os.environ.get("K", 6))


def timed(label, fn, pre=None):
    fn(False)
    eng.wait_download()
    torch.cuda.synchronize()
    ops.profile_enable(True)
    t0 = time.perf_counter()
    if pre:
        pre()
    for k in range(K):
        fn(k + 1 < K)
    eng.wait_download()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K * 1e3
    ops.profile_enable(False)
    prof = ops.profile_results()
    ksum = sum(v[0] for v in prof.values()) / K
    ing = prof.get("ingest_normalize", (0, 0, 0))[0] / K
    print(f"{label:12s} {dt:8.2f} ms/step   kernels {ksum:8.2f} ms (ingest {ing:5.2f})   gap {dt - ksum:6.2f}")


with torch.no_grad():
    timed("resident", lambda more: eng.run(raster, ids, R, row_offset=i0))
    timed("+download", lambda more: eng.run(raster, ids, R, row_offset=i0, map_out=host_map))
    timed("raw_device", lambda more: eng.run(dev_raw, ids, R, row_offset=i0))
    timed("raw_dev+dl", lambda more: eng.run(dev_raw, ids, R, row_offset=i0, map_out=host_map))

    def host_step(dl):
        def f(more):
            eng.run(host_raw, ids, R, row_offset=i0, map_out=host_map if dl else None)
            if more:
                eng.prefetch(host_raw, i0)
        return f
    timed("raw_host", host_step(False), pre=lambda: eng.prefetch(host_raw, i0))
    timed("e2e", host_step(True), pre=lambda: eng.prefetch(host_raw, i0))
