"""Where the fine-tuning step's (N4) and the census train step's wall time goes: torch.profiler over a few steps (development tool)."""
import os
import sys
import time
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import popcorn_b200 as pb  # noqa: E402
from popcorn_b200 import synthetic as sy  # noqa: E402

dev = torch.device("cuda", 0)
sd, _ = bench.bench_weights()
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    model = pb.POPCORN(6, occupancymodel=True, pretrained=False, biasinit=0.9407, sentinelbuildings=True, device=dev)
model.load_state_dict(sd)
model.train()
x, admin, cidx, y = bench._train_batch(dev)
for mode in ("head_only", "finetune"):
    ft = mode == "finetune"
    params = [p for n, p in model.named_parameters() if n.startswith("head.") or (ft and n.startswith("unetmodel."))]
    opt = torch.optim.Adam(params, lr=1e-5)

    def step():
        inp = {"input": x, "admin_mask": admin, "census_idx": cidx}
        out = model(inp, train=True, padding=False, encoder_no_grad=not ft, unet_no_grad=not ft, sparse=True)
        sy.census_loss(out, y).backward()
        torch.nn.utils.clip_grad_norm_([p for p in params if p.requires_grad], 0.01)
        opt.step()
        opt.zero_grad()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    print(f"== {mode}: {1e3 * (time.perf_counter() - t0) / 5:.2f} ms per step")
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=22, max_name_column_width=60))
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=14, max_name_column_width=60))
