"""The reference algorithm as stock PyTorch/cuDNN on ONE B200 — the "before" of SURVEY.md §8(d), next to bench.py's CPU leg.

Development/measurement tool (not part of the product, not used by bench.py): it runs the oracle's torch restatement of
POPCORN.forward — the same ATen/cuDNN ops the reference model executes (conv2d, batch_norm(eval), relu, max_pool2d,
conv_transpose2d, reflect pad, 1x1 convs, index/sum) — on CUDA tensors over reference-sized 2048^2 tiles, fp32 with TF32
off (utils/utils.py:57-58 sets cudnn.deterministic; the reference never enables TF32) and, for information, with TF32 on.

    python tools/torch_gpu_baseline.py            # prints one JSON line; unique px per tile = 1792^2 (centre write-back)
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import popcorn_oracle as po  # noqa: E402


def run(tf32: bool, tiles: int, H: int = 2048, W: int = 2048):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    dev = torch.device("cuda", 0)
    sd = {k: v.to(dev) for k, v in po.random_state_dict(seed=1600).items()}
    ids = po.synthetic_regions(H, W, 40).to(dev)
    x = po.synthetic_input(H, W, seed=1610).to(dev)

    def one():
        with torch.no_grad():
            out = po.forward(sd, {"input": x}, padding=False)
            centre = out["popdensemap"][0][128:-128, 128:-128]
            return torch.zeros(41, dtype=torch.float64, device=dev).index_add_(
                0, ids[128:-128, 128:-128].reshape(-1).long(), centre.reshape(-1).double())

    for _ in range(2):
        one()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(tiles):
        s = one()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return tiles * 1792 * 1792 / dt, float(s.sum()), torch.cuda.max_memory_allocated() / 2 ** 30


if __name__ == "__main__":
    assert torch.cuda.is_available(), "needs a CUDA device"
    n = int(os.environ.get("TILES", 5))
    v32, chk32, gib = run(False, n)
    vtf, chktf, _ = run(True, n)
    print(json.dumps({"what": "oracle restatement of POPCORN.forward on stock PyTorch/cuDNN, 2048^2 tiles, 1 GPU",
                      "torch": torch.__version__, "px_per_s_fp32": v32, "px_per_s_tf32": vtf, "tiles": n,
                      "check_fp32": chk32, "check_tf32": chktf, "peak_mem_gib": gib}))
