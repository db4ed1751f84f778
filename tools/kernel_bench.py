"""Per-kernel timings on one B200 (CUDA events, inputs larger than L2, >=3 warm-ups).  Development tool."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from popcorn_b200 import ops, weights  # noqa: E402
from oracle import popcorn_oracle as po  # noqa: E402


def timeit(fn, warm=3, it=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(it + 1)]
    ev[0].record()
    for i in range(it):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(it))
    return ts[len(ts) // 2]


def main():
    H = W = int(os.environ.get("KB_SIZE", 2048))
    sd = po.random_state_dict()
    x = torch.randn(1, 6, H, W, device="cuda")
    res = {"size": H, "f32x2": os.environ.get("POPCORN_CONV_F32X2", "0")}
    for copy, mode, pads, flop in (("unetmodel", 0, (0, 0, 0, 0), 18080), ("building_extractor", 1, (14,) * 4, 18096)):
        pack = weights.pack_dda(sd, copy).cuda()
        ms = timeit(lambda: ops.dda_forward(pack, x, pads, mode))
        res[f"dda_{copy}_ms"] = ms
        res[f"dda_{copy}_tflops"] = flop * H * W / ms / 1e9
        res[f"dda_{copy}_mpx_s"] = H * W / ms / 1e3
    feats = torch.randn(1, 16, H, W, device="cuda")
    bu = torch.rand(1, 1, H, W, device="cuda")
    hp = weights.pack_head(sd).cuda()
    ids = po.synthetic_regions(H, W, 400).cuda()[None].contiguous()
    sums = torch.zeros(401, dtype=torch.float64, device="cuda")
    ms = timeit(lambda: ops.head_dense_forward(hp, feats, bu, ids, None, sums))
    res["head_dense_ms"] = ms
    res["head_dense_tflops"] = 18688 * H * W / ms / 1e9
    res["head_dense_gbs"] = 80 * H * W / ms / 1e6
    hp_tc = weights.pack_head_tc(sd).cuda()
    ms = timeit(lambda: ops.head_dense_forward(hp_tc, feats, bu, ids, None, sums, tc=True))
    res["head_tc_ms"] = ms
    res["head_tc_tflops"] = 18688 * H * W / ms / 1e9
    res["head_tc_gbs"] = 80 * H * W / ms / 1e6
    n = 1 << 28
    d = torch.rand(n, device="cuda")
    big_ids = po.synthetic_regions(16384, 16384, 400).cuda().reshape(-1).contiguous()
    ms = timeit(lambda: ops.region_sum(d, big_ids, 401))
    res["region_sum_ms"] = ms
    res["region_sum_gbs"] = 8 * n / ms / 1e6
    from popcorn_b200 import _lib
    o = torch.zeros(4, device="cuda")
    for x2 in (0, 1):
        iters = 4096
        nthr = [0]
        def run():
            nthr[0] = _lib.lib().pc_test_fma_peak(x2, iters, 8, o.data_ptr(), torch.cuda.current_stream().cuda_stream)
        ms = timeit(run)
        res[f"fma_peak_x2_{x2}_tflops"] = 2 * 32 * iters * nthr[0] / ms / 1e9
    print(json.dumps(res))


if __name__ == "__main__":
    main()
