#!/usr/bin/env python
"""bench.py — POPCORN country inference throughput on B200 (contract: see the task statement / DESIGN.md §measurement).

    python bench.py --gpus 1 --steps 5 --warmup 3                       # this repo's CUDA path
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                                # the reference algorithm on host cores

One step = one full pass of the hot path over a synthetic country raster: tiled DDA builtup + feature passes,
occupancy head, centre-masked accumulation, finalisation, census region sums, all-reduce of the R sums.
N=1 : Rwanda-shaped 15104 x 17216 (BASELINE.json configs[1]).   N>1 : Uganda-shaped columns (47952) with
50048*N/8 rows, i.e. configs[3] at N=8 and the same per-GPU slab at N=2,4 (weak scaling, rows sharded).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
import warnings

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "10m_pixels_per_sec_country_inference"
UNIT = "pixels/s"
R_REGIONS = 400
FLOP_PER_PX_HEAD = 18688          # 9 344 MAC per pixel, dense head
BYTES_PER_PX_HEAD = 64 + 4 + 8    # 16 feature planes + builtup in, dens + scale out


def workload(n_gpus: int):
    if n_gpus <= 1:
        return 15104, 17216, "rwanda_shaped_15104x17216_tiled_inference_dense_head_region_sums_R400"
    H = 50048 * n_gpus // 8
    return H, 47952, f"uganda_shaped_{H}x47952_rows_sharded_over_{n_gpus}_gpus_R400"


# ---------------------------------------------------------------------------------------------------
# synthetic data, generated on the device slab by slab
# ---------------------------------------------------------------------------------------------------
def synth_raster_slab(rows: int, W: int, row0: int, device, seed: int = 1610) -> torch.Tensor:
    """[6, rows, W] normalised fp32, channel order [R,G,B,NIR,VV,VH]; low-frequency structure + noise."""
    g = torch.Generator(device=device).manual_seed(seed + row0)
    out = torch.empty(6, rows, W, dtype=torch.float32, device=device)
    for c in range(6):
        coarse = torch.randn(1, 1, rows // 64 + 2, W // 64 + 2, generator=g, device=device)
        low = torch.nn.functional.interpolate(coarse, size=(rows, W), mode="bilinear", align_corners=True)[0, 0]
        out[c] = 0.6 * low + 0.8 * torch.randn(rows, W, generator=g, device=device)
        del low
    return out


def synth_ids_slab(H: int, W: int, R: int, lo: int, hi: int, device, seed: int = 7) -> torch.Tensor:
    """int32 [hi-lo, W]: 0 = background frame, 1..R Voronoi cells of R seeded centres over the whole raster."""
    g = torch.Generator().manual_seed(seed)
    cy = (torch.rand(R, generator=g) * H).to(device)
    cx = (torch.rand(R, generator=g) * W).to(device)
    step = 16
    ys = torch.arange(lo // step * step, hi + step, step, device=device, dtype=torch.float32)
    xs = torch.arange(0, W + step, step, device=device, dtype=torch.float32)
    best = torch.full((len(ys), len(xs)), float("inf"), device=device)
    ids = torch.zeros(len(ys), len(xs), dtype=torch.int32, device=device)
    for r in range(R):
        d = (ys[:, None] - cy[r]) ** 2 + (xs[None, :] - cx[r]) ** 2
        upd = d < best
        best = torch.where(upd, d, best)
        ids[upd] = r + 1
    full = ids.repeat_interleave(step, 0).repeat_interleave(step, 1)
    off = lo - lo // step * step
    full = full[off: off + (hi - lo), :W].contiguous()
    m = 128
    if lo < m:
        full[: m - lo] = 0
    if hi > H - m:
        full[max(0, H - m - lo):] = 0
    full[:, :m] = 0
    full[:, W - m:] = 0
    return full



# per-kernel algorithmic work: FLOP and HBM bytes per processed pixel of that launch
def _conv_work(name):
    ca, cb, co, epi = name[name.index("<") + 1:-1].split(",")
    cin, co = int(ca) + int(cb), int(co)
    flop = 2 * 9 * cin * co
    # convt: the Up block's ConvTranspose2d 2x2 rides in the epilogue: 4 output pixels x co channels written per input pixel
    out_b = {"store": co * 4, "pool": co * 4 + co, "dot": 4 + 4, "convt": 4 * co * 4}[epi]
    if epi == "convt":
        flop += 2 * 4 * co * co
    return flop, cin * 4 + out_b


FLOP_PER_SEL_PX_HEAD_BWD = 2 * (9280 + 8256 + 9280)   # recomputed forward + dgrad (64x64 x2 + 64) + wgrad, per selected pixel


def ncu_traffic_per_px():
    """dram__bytes_read.sum + dram__bytes_write.sum per processed pixel and kernel, from the committed `ncu --set full` capture of
    THIS command's launches (profiles/r2_ncu_bench_traffic.json, written by tools/ncu_traffic.py from the ncu CSV); {} if absent."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_bench_traffic.json")
    try:
        return {k: v["dram_bytes_per_px"] for k, v in json.load(open(p))["kernels"].items()}
    except Exception:
        return {}


def kernel_table(prof, ms_total, hbm_peak, tensor_peak, fp32_peak=None, frac_multi=0.0, maps=4):
    """prof = {name: (ms, launches, units)} from ops.profile_results().  `bound`: "hbm" | "tensor" as in the contract, plus "fp32" for the
    kernels whose ceiling is the FP32 FFMA2 pipe (measured by pc_test_fma_peak), which is neither."""
    traffic_px = ncu_traffic_per_px()
    from popcorn_b200 import weights as _w
    f16 = _w.tc_operand_format() == 1
    # fp32-accurate tensor-core arithmetic = three split products per algorithmic MAC: fp16 halves (kind::f16) run at the bf16 rate the
    # peak was measured with, TF32 halves (kind::tf32) at half of it
    slots = 3 if f16 else 6
    split = ("3 x fp16-half products (kind::f16, x = hi + lo with 11 + 11 significand bits)" if f16
             else "3xTF32 (kind::tf32 at half the bf16 rate)")
    rows = []
    for name, (ms, n, units) in prof.items():
        if ms <= 0:
            continue
        if name.startswith("conv3x3_tc<"):
            flop, byts = _conv_work(name)
            bound = "hbm"
            note = (f"tcgen05 implicit GEMM, {split}: tensor ceiling = peak/{slots} = {tensor_peak / slots:.0f} TFLOP/s; "
                    "the layer moves (Cin+Cout)*4 B per pixel and is HBM-bound")
        elif name.startswith("conv3x3<"):
            flop, byts = _conv_work(name)
            bound, note = "hbm", "fp32 FFMA2 stencil (first layer: reflect padding + channel remap, Cin 2|4): 12-48 B and 144-288 MAC per pixel"
        elif name.startswith("convt2x2<"):
            c = int(name[len("convt2x2<"):-1])
            flop, byts, bound, note = 8 * c * c, 5 * c * 4, "hbm", "units = low-res pixels"
        elif name.startswith("head_tc"):
            flop, byts, bound = FLOP_PER_PX_HEAD, BYTES_PER_PX_HEAD, "tensor"
            note = f"tcgen05, activations in TMEM, {split}: ceiling = peak/{slots}"
        elif name.startswith("head_forward_simt"):
            flop, byts, bound, note = FLOP_PER_PX_HEAD, BYTES_PER_PX_HEAD, "fp32", "fp32 SIMT head"
        elif name == "head_backward":
            flop, byts, bound = FLOP_PER_SEL_PX_HEAD_BWD, 64 + 4 + 4 + 4, "fp32"
            note = "units = selected pixels; forward recomputed in shared memory + dgrad + wgrad on the FP32 pipe (csrc/head_bwd.cu)"
        elif name.startswith("compact_") or name == "sparse_mask_compact":
            flop, byts, bound, note = 2, 9, "hbm", "mask = (builtup>0)&(admin==idx)|grid, row-major compaction: 9 B/px + 4 B per selected"
        elif name == "ingest_normalize":
            flop, byts, bound, note = 12, 16 + 24, "hbm", "uint16 S2 x4 + float32 S1 x2 read, 6 fp32 planes written"
        elif name == "region_sum":
            flop, byts, bound, note = 1, 8, "hbm", "dens + id read once"
        elif name == "accumulate":
            flop, byts, bound, note = 4, 8 + maps * 8 + 4, "hbm", f"tile dens+scale read, {maps} maps + count read-modify-write"
        elif name == "finalize":
            flop, byts, bound = 10, 2 + maps * 8 * frac_multi, "hbm"
            note = f"int16 count read for every pixel; {maps} maps read+written only where count > 1 ({100 * frac_multi:.1f} % of the pixels here)"
        else:
            flop, byts, bound, note = 0, 0, "hbm", ""
        sec = ms * 1e-3
        tfl = flop * units / sec / 1e12
        gbs = byts * units / sec / 1e9
        if bound == "hbm":
            ach, peak, unit = gbs, hbm_peak, "GB/s"
        elif bound == "fp32":
            ach, peak, unit = tfl, fp32_peak, "TFLOP/s"
        else:
            ach, peak, unit = tfl, tensor_peak, "TFLOP/s"
        tpp = traffic_px.get(name)
        rows.append({"kernel": name, "ms": ms, "launches": n, "pixels": units, "share_of_step": ms / ms_total if ms_total else None,
                     "tflops": tfl, "gbs": gbs, "bound": bound, "achieved": ach, "peak": peak, "unit": unit,
                     "frac": ach / peak if peak else None, "algorithmic_bytes": byts * units / n,
                     "traffic": None if tpp is None else tpp * units / n, "note": note})
        if name.startswith("head_tc"):
            rows[-1]["ceiling_split3"] = tensor_peak / slots
            rows[-1]["frac_of_split3_ceiling"] = tfl / (tensor_peak / slots) if tensor_peak else None
            rows[-1]["operands"] = "fp16 hi/lo" if f16 else "tf32 hi/lo"
    rows.sort(key=lambda r: -r["ms"])
    return rows


def roofline_of(kernels, peak_src):
    dom = max(kernels, key=lambda k: k["ms"]) if kernels else None
    if dom is None:
        return None
    r = {"kernel": dom["kernel"], "bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"], "unit": dom["unit"],
         "frac": dom["frac"], "traffic": dom.get("traffic"), "algorithmic_bytes": dom.get("algorithmic_bytes"),
         "traffic_source": "ncu --set full capture of this command's launches (profiles/r2_ncu_bench_traffic.json)" if dom.get("traffic") else None,
         "launches": dom["launches"], "avg_launch_ms": dom["ms"] / dom["launches"], "share_of_step": dom["share_of_step"],
         "note": dom["note"], "peak_source": peak_src}
    for k in ("ceiling_split3", "frac_of_split3_ceiling", "operands"):
        if k in dom:
            r[k] = dom[k]
    return r


def measure_fp32_peak(dev):
    """Register-resident FFMA2 loop (pc_test_fma_peak): the practical ceiling of the FP32 pipe, TFLOP/s."""
    from popcorn_b200 import _lib
    o = torch.zeros(4, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    iters = 4096
    for _ in range(2):
        nthr = _lib.lib().pc_test_fma_peak(1, iters, 8, o.data_ptr(), st)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    nthr = _lib.lib().pc_test_fma_peak(1, iters, 8, o.data_ptr(), st)
    b.record()
    torch.cuda.synchronize()
    return 2 * 32 * iters * nthr / (a.elapsed_time(b) * 1e-3) / 1e12


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        res = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return res
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            res = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return res


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


# ---------------------------------------------------------------------------------------------------
# CPU baseline = the oracle port of the reference algorithm on the host cores (bounded sample)
# ---------------------------------------------------------------------------------------------------
def bench_weights():
    """BASELINE config 1's weight recipe: `unetmodel` re-initialised by the reference's own init code at seed 1600 (kaiming fan_out, BN
    gamma 1 / beta 0, running stats of the checkpoint), default-init head with bias 0.9407, `building_extractor` = the pretrained DDA
    checkpoint.  That state_dict was produced by the UNMODIFIED reference (oracle/make_golden.py) and is committed as a fixture
    (tests/golden/state_dict.npz, 324 keys); without it (stripped checkout) the synthetic generator of popcorn_b200.synthetic is used."""
    p = os.path.join(ROOT, "tests", "golden", "state_dict.npz")
    if os.path.isfile(p) and os.environ.get("POPCORN_BENCH_WEIGHTS", "reference") != "synthetic":
        import numpy as np
        return {k: torch.from_numpy(np.asarray(v)) for k, v in np.load(p).items()}, \
            "config-1 recipe from the unmodified reference (tests/golden/state_dict.npz: pretrained DDA checkpoint in building_extractor, " \
            "reference re-init of unetmodel + head at seed 1600)"
    from popcorn_b200 import synthetic as sy
    return sy.random_state_dict(seed=1600), "random-init DDA x2 + head (popcorn_b200.synthetic.random_state_dict seed 1600)"


def cpu_reference_tiles(n_tiles: int, sd=None, keep=None):
    """Runs n_tiles reference-sized (2048^2) tiles through the oracle + census sums; returns (seconds, unique px).
    keep: optional dict that receives tile 0's input, centre density and region sums (for the GPU-vs-oracle check of the bench)."""
    from oracle import popcorn_oracle as po
    torch.set_num_threads(os.cpu_count() or 1)
    sd = sd or bench_weights()[0]
    ids = po.synthetic_regions(2048, 2048, 40)
    xs = [po.synthetic_input(2048, 2048, seed=1610 + i) for i in range(n_tiles)]   # untimed: data generation
    t0 = time.perf_counter()
    for i, x in enumerate(xs):
        with torch.no_grad():
            out = po.forward(sd, {"input": x}, padding=False)
        centre = out["popdensemap"][0][128:-128, 128:-128]
        sums = po.region_sums(centre, ids[128:-128, 128:-128], 41)
        if keep is not None and i == 0:
            keep.update(x=x, centre=centre.clone(), sums=sums.clone(), ids=ids)
    dt = time.perf_counter() - t0
    return dt, n_tiles * 1792 * 1792


def gpu_reference_tiles(n_tiles: int, sd, dev, tf32: bool):
    """The SAME oracle restatement (the ATen/cuDNN ops the reference model executes) on this GPU through stock PyTorch: the honest
    "before" of SURVEY.md §8(d).  fp32 with TF32 off is the reference's own setting (utils/utils.py:57-58 never enables TF32)."""
    from oracle import popcorn_oracle as po
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    try:
        sdd = {k: v.to(dev) for k, v in sd.items()}
        ids = po.synthetic_regions(2048, 2048, 40).to(dev)
        x = po.synthetic_input(2048, 2048, seed=1610).to(dev)

        def one():
            with torch.no_grad():
                out = po.forward(sdd, {"input": x}, padding=False)
                centre = out["popdensemap"][0][128:-128, 128:-128]
                return torch.zeros(41, dtype=torch.float64, device=dev).index_add_(
                    0, ids[128:-128, 128:-128].reshape(-1).long(), centre.reshape(-1).double())
        for _ in range(2):
            one()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_tiles):
            one()
        torch.cuda.synchronize()
        return n_tiles * 1792 * 1792 / (time.perf_counter() - t0)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    H, W, name = workload(args.gpus)
    sd, wname = bench_weights()
    torch.set_num_threads(os.cpu_count() or 1)
    # warm-up is bounded to one tile: every further warm-up tile costs seconds of host time and changes nothing
    for _ in range(min(args.warmup, 1)):
        cpu_reference_tiles(1, sd)
    times = []
    px = 0
    for _ in range(args.steps):
        dt, p = cpu_reference_tiles(1, sd)
        times.append(dt)
        px += p
    total = sum(times)
    value = px / total
    cores = torch.get_num_threads()
    sample = "per step: one 2048x2048 reference tile (oracle port of POPCORN.forward, fp32) + census sums; unique px = centre 1792^2"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "sampled": sample, "weights": wname},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def _train_batch(device):
    B, H, W = 2, 896, 960
    x = synth_raster_slab(B * H, W, 12345, device).view(6, B, H, W).permute(1, 0, 2, 3).contiguous()
    yy, xx = torch.meshgrid(torch.arange(H, device=device), torch.arange(W, device=device), indexing="ij")
    admin = torch.zeros(B, H, W, device=device)
    admin[0][((yy - 448) / 400.0) ** 2 + ((xx - 480) / 420.0) ** 2 < 1] = 17.0
    admin[1][((yy - 430) / 380.0) ** 2 + ((xx - 500) / 400.0) ** 2 < 1] = 5.0
    cidx = torch.tensor([17, 5], device=device)
    y = torch.tensor([3500.0, 12000.0], device=device)
    return x, admin, cidx, y


def oracle_train_step_ms(sd, batch, device, iters: int, tf32: bool = False):
    """The reference algorithm's census train step (run_train.py:201-238 with unet_no_grad: two frozen DDA passes, sparse head, log-L1
    + scale regulariser, backward, clip 0.01, Adam) as stock PyTorch autograd through the oracle restatement, on `device` (cpu | cuda)."""
    from oracle import popcorn_oracle as po
    x, admin, cidx, y = (t.to(device) for t in batch)
    sdd = {k: v.to(device) for k, v in sd.items()}
    params = [sdd[k].requires_grad_(True) for k in sdd if k.startswith("head.")]
    opt = torch.optim.Adam(params, lr=1e-4)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    times = []
    try:
        for it in range(iters + 1):
            if x.is_cuda:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = po.forward(sdd, {"input": x, "admin_mask": admin, "census_idx": cidx}, padding=False, sparse=True, unet_no_grad=True)
            po.train_loss(out, y).backward()
            torch.nn.utils.clip_grad_norm_(params, 0.01)
            opt.step()
            opt.zero_grad()
            if x.is_cuda:
                torch.cuda.synchronize()
            if it >= 1:
                times.append(1e3 * (time.perf_counter() - t0))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return sorted(times)[len(times) // 2]


def train_step_section(model, sd, device, peaks, fp32_peak, cpu_baseline: bool, iters: int = 5):
    """BASELINE config 3 / metric (ii): census-supervised step B=2, 896x960, sparse head, log-L1 loss, clip 0.01, Adam (run_train.py),
    with its own per-kernel table, roofline (dominant kernel of the step), CPU baseline and stock-PyTorch GPU baseline."""
    from popcorn_b200 import ops
    from popcorn_b200 import synthetic as sy
    hbm_peak, bf16_peak, bf16_sust, peak_src = peaks
    batch = _train_batch(device)
    x, admin, cidx, y = batch
    B, _, H, W = x.shape
    model.train()
    params = [p for n, p in model.named_parameters() if n.startswith("head.")]
    opt = torch.optim.Adam(params, lr=1e-4)
    times = []
    n_sel = 0
    prof, ev_ms = {}, 0.0
    for it in range(iters + 3):
        profiled = it == iters + 2          # the last iteration runs with the per-launch CUDA events on (not part of the median)
        if profiled:
            ops.profile_enable(True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        inp = {"input": x, "admin_mask": admin, "census_idx": cidx}
        out = model(inp, train=True, padding=False, encoder_no_grad=True, unet_no_grad=True, sparse=True)
        loss = sy.census_loss(out, y)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 0.01)
        opt.step()
        opt.zero_grad()
        e1.record()
        torch.cuda.synchronize()
        if profiled:
            ops.profile_enable(False)
            prof, ev_ms = ops.profile_results(), e0.elapsed_time(e1)
        elif it >= 2:
            times.append(1e3 * (time.perf_counter() - t0))
        n_sel = int(out["scale"].numel())
    ms = sorted(times)[len(times) // 2]
    kernels = kernel_table(prof, ev_ms, hbm_peak, bf16_sust, fp32_peak)
    res = {"metric": "census_train_step_ms", "value": ms, "ms": ms, "unit": "ms", "higher_is_better": False,
           "config": f"B=2 x {H}x{W}, sparse head on {n_sel} px, unet_no_grad, log-L1 + scale reg, clip 0.01, Adam",
           "includes": "2 frozen DDA passes + mask compaction + sparse head fwd + loss + head bwd + clip + Adam (host-timed, one sync per step "
                       "as in the reference's n = mask.sum())",
           "px_per_step": B * H * W, "selected_px": n_sel, "kernels": kernels, "roofline": roofline_of(kernels, peak_src),
           "kernel_ms_sum": sum(k["ms"] for k in kernels), "profiled_step_ms": ev_ms}
    try:   # stock PyTorch / cuDNN on this GPU (fp32, TF32 off = the reference's setting; TF32 on for information)
        g32 = oracle_train_step_ms(sd, batch, device, 3, tf32=False)
        gtf = oracle_train_step_ms(sd, batch, device, 3, tf32=True)
        res["gpu_baseline"] = {"ms": g32, "ms_tf32": gtf, "kind": "port", "speedup_over_fp32": g32 / ms,
                               "what": "oracle restatement of the reference step as stock PyTorch autograd / cuDNN on the same B200"}
    except Exception as ex:
        res["gpu_baseline"] = {"error": repr(ex)[:200]}
    if cpu_baseline:
        try:
            torch.set_num_threads(os.cpu_count() or 1)
            c = oracle_train_step_ms(sd, batch, "cpu", 2)
            res["cpu_baseline"] = {"value": c, "unit": "ms", "cores": torch.get_num_threads(), "kind": "port",
                                   "sample": "the full step (B=2 x 896x960), median of 2 after 1 warm-up, oracle port on all host threads"}
        except Exception as ex:
            res["cpu_baseline"] = {"error": repr(ex)[:200]}
    # the reference's default for batches < 9 M px: unetmodel is fine-tuned as well (run_train.py:191-202)
    try:
        params_ft = [p for n, p in model.named_parameters() if n.startswith("head.") or n.startswith("unetmodel.")]
        opt2 = torch.optim.Adam([p for p in params_ft], lr=1e-5)
        ft = []
        for it in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            inp = {"input": x, "admin_mask": admin, "census_idx": cidx}
            out = model(inp, train=True, padding=False, encoder_no_grad=False, unet_no_grad=False, sparse=True)
            sy.census_loss(out, y).backward()
            torch.nn.utils.clip_grad_norm_([p for p in params_ft if p.requires_grad], 0.01)
            opt2.step()
            opt2.zero_grad()
            torch.cuda.synchronize()
            if it >= 1:
                ft.append(1e3 * (time.perf_counter() - t0))
        res["finetune_ms"] = sorted(ft)[len(ft) // 2]
        res["finetune_includes"] = "builtup pass + layer-by-layer unetmodel forward (activations kept) + sparse head fwd/bwd + full UNet backward + clip + Adam"
    except Exception as ex:
        res["finetune_error"] = repr(ex)[:200]
    model.load_state_dict(sd)      # undo the updates: everything after this uses the benchmark weights again
    model.eval()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="popcorn_b200", choices=["popcorn_b200", "reference"])
    ap.add_argument("--rows-per-strip", type=int, default=3, help="tile-rows per merged window (first and last strip: 1, see --edge-strip-rows)")
    ap.add_argument("--edge-strip-rows", type=int, default=1, help="tile-rows of the first and of the last strip of a rank")
    ap.add_argument("--no-merge", action="store_true", help="run the reference's 2048^2 tile grid tile by tile")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-train", action="store_true")
    ap.add_argument("--skip-timeseries", action="store_true")
    ap.add_argument("--no-balance", action="store_true",
                    help="N>1: shard whole row strips instead of 256-row units (country.plan_balanced_shards is the default)")
    ap.add_argument("--per-window-upload", action="store_true",
                    help="e2e: upload every window separately (halo rows cross PCIe twice) instead of each raw row once")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not bind the rank to its GPU's NUMA node")
    ap.add_argument("--no-e2e-pipeline", action="store_true", help="e2e: synchronise after every raster instead of prefetching the next one")
    ap.add_argument("--skip-ensemble", action="store_true")
    ap.add_argument("--skip-gpu-baseline", action="store_true")
    ap.add_argument("--skip-alone", action="store_true", help="N>1: skip the one-rank-alone run of rank 0's slab")
    ap.add_argument("--height", type=int, default=0, help="override raster rows (debug)")
    ap.add_argument("--width", type=int, default=0, help="override raster cols (debug)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: popcorn_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import popcorn_b200 as pb
    from popcorn_b200 import country as ct
    from popcorn_b200 import numa, ops

    # host threads (and the pinned staging buffers they first-touch) on the NUMA node of this rank's GPU
    numa_info = numa.bind_to_gpu_node(local) if not args.no_numa_bind else {"bound": False}

    H, W, name = workload(world)
    if args.height and args.width:
        H, W, name = args.height, args.width, f"debug_{args.height}x{args.width}"
    sd, weights_name = bench_weights()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = pb.POPCORN(6, occupancymodel=True, pretrained=False, biasinit=0.9407, sentinelbuildings=True, device=dev)
    model.load_state_dict(sd)
    model.eval()

    balance = not args.no_balance
    upload_once = not args.per_window_upload
    eng = ct.CountryEngine([model], H, W, merge=not args.no_merge, rows_per_strip=args.rows_per_strip, rank=rank, world=world,
                           first_strip_rows=args.edge_strip_rows, last_strip_rows=args.edge_strip_rows, balance=balance,
                           upload_once=upload_once)     # short first / last strips: compute starts after a small upload, little map left to ship at the end
    i0, i1 = eng.in_rows
    lo, hi = eng.out_rows
    raster = synth_raster_slab(i1 - i0, W, i0, dev)
    ids = synth_ids_slab(H, W, R_REGIONS, lo, hi, dev)
    R = R_REGIONS + 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        with torch.no_grad():
            return eng.run(raster, ids, R, row_offset=i0)

    for _ in range(args.warmup):
        out = step_device()
    barrier()
    ops.launch_count(reset=True)
    ops.profile_enable(True)        # CUDA events around every launch of the library, on the launch stream
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out = step_device()
    ev1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ops.profile_enable(False)
    launches = ops.launch_count(reset=True)
    prof = ops.profile_results()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = H * W / (ms_step * 1e-3)
    sums_check = float(out["sums"].sum().item())          # all-reduced census sums (every id incl. background 0)
    mt = out["map"].double().sum().reshape(1)               # this rank's rows of the map -> all ranks
    if world > 1:
        dist.all_reduce(mt)
    map_total = float(mt.item())
    frac_multi = float((out["count"] > 1).float().mean().item()) if out["count"].numel() else 0.0

    peaks = measured_peaks()
    hbm_peak, bf16_peak, bf16_sust, peak_src = peaks
    fp32_peak = measure_fp32_peak(dev)
    kernels = kernel_table(prof, ms_total, hbm_peak, bf16_sust, fp32_peak, frac_multi)
    roofline = roofline_of(kernels, peak_src)

    # ---- N > 1: the same per-GPU slab on ONE rank alone (the other ranks idle at the barrier): the apples-to-apples N = 1 point of
    #      the weak-scaling series (the N = 1 line of the contract is the Rwanda-shaped raster, another shape)
    alone = None
    if world > 1 and not args.skip_alone:
        barrier()
        if rank == 0:
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(2):
                with torch.no_grad():
                    eng.run(raster, ids, R, row_offset=i0, reduce=False)
            a1.record()
            torch.cuda.synchronize()
            px_rank = (hi - lo) * W
            alone = {"ms": a0.elapsed_time(a1) / 2, "px": px_rank, "value": px_rank / (a0.elapsed_time(a1) / 2 * 1e-3),
                     "note": "rank 0's slab with the other ranks idle, no collective: per-GPU rate without neighbours"}
        barrier()

    # ---- end-to-end: pinned host raster in, pinned host map + sums out, copies inside the timed region ----
    e2e = None
    if not args.skip_e2e:
        # host rasters in the on-disk dtypes the reference reads (uint16 S2 in file band order B,G,R,NIR + float32 S1 dB):
        # de-normalise the synthetic raster, quantise S2, pin.  RawRaster uploads 16 B/px and normalises on the device.
        st2, st1 = ops.DATASET_STATS["sen2springNIR"], ops.DATASET_STATS["sen1"]
        rows = raster.shape[1]
        host_s2 = torch.empty(4, rows, W, dtype=torch.uint16, pin_memory=True)
        host_s1 = torch.empty(2, rows, W, dtype=torch.float32, pin_memory=True)
        for dst_plane, c in enumerate((2, 1, 0, 3)):        # file order B02,B03,B04,B08 <- R,G,B,NIR channels 2,1,0,3
            v = (raster[c] * st2["std"][c] + st2["mean"][c]).clamp_(0, 10000).round_()
            host_s2[dst_plane].copy_(v.to(torch.int32).to(torch.uint16))
            del v
        for c in range(2):
            host_s1[c].copy_(raster[4 + c] * st1["std"][c] + st1["mean"][c])
        host_raster = ct.RawRaster(host_s2, host_s1, ops.S2_FILE_TO_RGBN)
        host_map = torch.empty(hi - lo, W, dtype=torch.float32, pin_memory=True)
        del raster
        torch.cuda.empty_cache()

        pipelined = upload_once and not args.no_e2e_pipeline
        sums_host = [torch.empty(R, dtype=torch.float64, pin_memory=True) for _ in range(8)]
        k_slot = [0]

        def step_e2e(more: bool):
            """One raster through the public API: host bands in, host map + census sums out.  Pipelined (default): the NEXT raster's
            upload is queued (prefetch) before this step's result is read, and the tail of this step's map download overlaps the next
            step's first windows — how a stream of rasters (seasonal frames, successive countries) runs; every step's copies still
            happen inside the timed region, and the region ends with all downloads complete."""
            with torch.no_grad():
                o = eng.run(host_raster, ids, R, row_offset=i0, map_out=host_map)    # finished strips stream back during compute
                if pipelined and more:
                    eng.prefetch(host_raster, i0)
                if pipelined:      # the step's result: census sums -> pinned host memory, asynchronously (read after the region's final sync)
                    s = sums_host[k_slot[0] % len(sums_host)]
                    k_slot[0] += 1
                    s.copy_(o["sums"], non_blocking=True)
                else:
                    s = o["sums"].cpu()
            if not pipelined:
                eng.wait_download()
                torch.cuda.synchronize()
            return s

        step_e2e(False)
        eng.wait_download()
        barrier()
        t0 = time.perf_counter()
        k_e2e = max(3, min(args.steps, 20))         # the un-overlapped first upload and last download are inside the region: ~80 ms / k_e2e per step
        if pipelined:
            eng.prefetch(host_raster, i0)
        for k_ in range(k_e2e):
            s_host = step_e2e(k_ + 1 < k_e2e)
        eng.wait_download()
        barrier()                                  # barrier() synchronises the device: every step's sums and map rows are on the host now
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        h2d = torch.tensor([float(getattr(eng, "h2d_bytes", 0))], dtype=torch.float64, device=dev)
        d2h = torch.tensor([float(host_map.numel() * 4 + R * 8)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(h2d); dist.all_reduce(d2h)
        # what the host <-> device links of this box deliver when every rank copies at once (1 GiB up and 256 MiB down per rank,
        # concurrently, pinned memory): the ceiling of any end-to-end number that moves 16 B/px up and 4 B/px down
        pcie = None
        try:
            nb_up, nb_dn = 1 << 30, 1 << 28
            hb_up = torch.empty(nb_up, dtype=torch.uint8, pin_memory=True)
            hb_dn = torch.empty(nb_dn, dtype=torch.uint8, pin_memory=True)
            db_up = torch.empty(nb_up, dtype=torch.uint8, device=dev)
            db_dn = torch.empty(nb_dn, dtype=torch.uint8, device=dev)
            s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
            res_bw = {}
            for mode in ("h2d", "d2h", "both"):
                barrier()
                t1 = time.perf_counter()
                for _ in range(2):
                    if mode in ("h2d", "both"):
                        with torch.cuda.stream(s_up):
                            db_up.copy_(hb_up, non_blocking=True)
                    if mode in ("d2h", "both"):
                        with torch.cuda.stream(s_dn):
                            hb_dn.copy_(db_dn, non_blocking=True)
                torch.cuda.synchronize()
                barrier()
                dtm = torch.tensor([time.perf_counter() - t1], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(dtm, op=dist.ReduceOp.MAX)
                sec = float(dtm.item())
                res_bw[mode] = {"h2d_gbs_all_ranks": (2 * nb_up * world / sec / 1e9) if mode != "d2h" else 0.0,
                                "d2h_gbs_all_ranks": (2 * nb_dn * world / sec / 1e9) if mode != "h2d" else 0.0}
            up, dn = res_bw["h2d"]["h2d_gbs_all_ranks"], res_bw["d2h"]["d2h_gbs_all_ranks"]
            upc, dnc = res_bw["both"]["h2d_gbs_all_ranks"], res_bw["both"]["d2h_gbs_all_ranks"]
            # a steady stream of rasters moves 16 B/px up and 4 B/px down AT THE SAME TIME: the bound uses the rates measured with both
            # directions active (the 4:1 byte ratio of the probe equals the path's)
            t_px = max(16.0 / (upc * 1e9), 4.0 / (dnc * 1e9))
            pcie = {"h2d_gbs_all_ranks_alone": up, "d2h_gbs_all_ranks_alone": dn, "concurrent": res_bw["both"],
                    "bound_px_per_s": 1.0 / t_px,
                    "note": "every rank copying at once; bound = 16 B/px up and 4 B/px down over these links with both directions active"}
            del hb_up, hb_dn, db_up, db_dn
        except Exception as ex:
            pcie = {"error": repr(ex)[:200]}
        e2e_value = H * W / (float(tt.item()) / k_e2e)
        if pcie and pcie.get("bound_px_per_s"):
            pcie["e2e_frac_of_bound"] = e2e_value / pcie["bound_px_per_s"]
        e2e = {"value": e2e_value, "unit": UNIT, "host_links": pcie, "h2d_bytes_per_step": int(h2d.item()),
               "d2h_bytes_per_step": int(d2h.item()), "steps": k_e2e,
               "api": "popcorn_b200.country.CountryEngine.run(RawRaster(pinned uint16 S2 + float32 S1), map_out=pinned host map) + sums.cpu()",
               "upload": "once_per_row" if upload_once else "per_window", "numa": numa_info,
               "pipelined": pipelined, "pipeline_note": "double-buffered input slabs: raster k+1 uploads while raster k computes; all copies inside the timed region" if pipelined else None,
               "host_input": "raw on-disk dtypes: S2 uint16 x4 (file band order) + S1 float32 x2 = 16 B/px; converted + normalised on the device"}
        del host_raster, host_s2, host_s1, host_map
    else:
        del raster
    torch.cuda.empty_cache()

    train = None
    cpu_base = None
    gpu_base = None
    fp32 = None
    check = {"sum_of_region_sums": sums_check, "map_total": map_total}
    ens = None
    if rank == 0:
        if not args.skip_train:
            try:
                train = train_step_section(model, sd, dev, peaks, fp32_peak, cpu_baseline=(world == 1 and not args.skip_cpu_baseline))
            except Exception as ex:   # the headline metric must still be printed
                train = {"error": repr(ex)[:300]}
        conv_ms = sum(k["ms"] for k in kernels if k["kernel"].startswith("conv3x3<"))
        conv_fl = sum(k["tflops"] * k["ms"] for k in kernels if k["kernel"].startswith("conv3x3<"))
        tc_ms = sum(k["ms"] for k in kernels if k["kernel"].startswith("conv3x3_tc<"))
        tc_fl = sum(k["tflops"] * k["ms"] for k in kernels if k["kernel"].startswith("conv3x3_tc<"))
        tc_gb = sum(k["gbs"] * k["ms"] for k in kernels if k["kernel"].startswith("conv3x3_tc<"))
        alg_b = sum(k["algorithmic_bytes"] * k["launches"] for k in kernels)
        trf = [k for k in kernels if k.get("traffic")]
        fp32 = {"peak_tflops_measured_ffma2": fp32_peak,
                "conv3x3_all_tflops": conv_fl / conv_ms if conv_ms else None,
                "conv3x3_all_frac_of_fp32_peak": conv_fl / conv_ms / fp32_peak if conv_ms and fp32_peak else None,
                "note": "register-resident FFMA2 loop (pc_test_fma_peak): the practical ceiling of the fp32 stencil kernels",
                "conv3x3_tc_all_tflops_algorithmic": tc_fl / tc_ms if tc_ms else None,
                "conv3x3_tc_all_gbs_algorithmic": tc_gb / tc_ms if tc_ms else None,
                "conv3x3_tc_all_frac_of_hbm_peak": tc_gb / tc_ms / hbm_peak if tc_ms else None,
                "path_bytes_per_unique_px_kernel_sum": alg_b / args.steps / (H * W / world) if kernels else None,
                "path_dram_bytes_per_unique_px_ncu": (sum(k["traffic"] * k["launches"] for k in trf) / args.steps / (H * W / world)) if trf else None}
        if world == 1 and not args.skip_gpu_baseline:
            # stock PyTorch / cuDNN on this very GPU: the honest "before" (SURVEY.md §8d), fp32 with TF32 off (the reference's setting) and on
            try:
                g32 = gpu_reference_tiles(3, sd, dev, tf32=False)
                gtf = gpu_reference_tiles(3, sd, dev, tf32=True)
                gpu_base = {"value": g32, "value_tf32": gtf, "unit": UNIT, "kind": "port", "speedup_over_fp32": value / g32,
                            "speedup_over_tf32": value / gtf,
                            "sample": "3 reference tiles of 2048x2048 (oracle restatement = the reference's ATen/cuDNN ops, stock PyTorch "
                                      f"{torch.__version__}) + census sums on the same B200; unique px = centre 1792^2 per tile"}
            except Exception as ex:
                gpu_base = {"error": repr(ex)[:200]}
            torch.cuda.empty_cache()
        if world == 1 and not args.skip_cpu_baseline:
            keep = {}
            dt, px = cpu_reference_tiles(4, sd, keep)
            cpu_base = {"value": px / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                        "sample": "4 reference tiles of 2048x2048 (oracle port of POPCORN.forward, fp32, all host threads) "
                                  "+ census sums; unique px = centre 1792^2 per tile"}
            # parity at bench size: the same 2048^2 tile through the product path (model API -> C-ABI), centre against the oracle's
            try:
                with torch.no_grad():
                    o = model({"input": keep["x"].to(dev)}, padding=False)
                    cg = o["popdensemap"][0][128:-128, 128:-128].contiguous()
                    sg = ops.region_sum(cg, keep["ids"][128:-128, 128:-128].contiguous().to(dev), 41).cpu()
                ref = keep["centre"].double()
                err = ((cg.cpu().double() - ref).abs() / torch.clamp(ref.abs(), min=1e-3 * float(ref.abs().max()))).max()
                nz = keep["sums"].abs() > 0
                serr = ((sg - keep["sums"]).abs()[nz] / keep["sums"].abs()[nz]).max()
                check.update(tile2048_density_max_rel=float(err), tile2048_region_sum_max_rel=float(serr),
                             tile2048_ok=bool(err < 1e-2 and serr < 1e-3),
                             tile2048_note="one 2048^2 tile of the cpu_baseline leg re-run through popcorn_b200 on the GPU; bars 1e-2 / 1e-3")
            except Exception as ex:
                check["tile2048_error"] = repr(ex)[:200]
            del keep

    # ---- 5-member ensemble (run_eval.py:108-115; README: seeds 1600-1604): members share the builtup pass ----
    if world == 1 and not args.skip_ensemble:
        try:
            members = [model]
            for k in range(1, 5):
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    mk = pb.POPCORN(6, occupancymodel=True, pretrained=False, biasinit=0.9407, sentinelbuildings=True, device=dev)
                sdk = dict(sd)
                g = torch.Generator().manual_seed(1600 + k)       # fine-tuned parts differ per member; building_extractor is shared
                for key, v in sd.items():
                    if (key.startswith("unetmodel.") or key.startswith("head.")) and v.is_floating_point() and "running" not in key:
                        sdk[key] = v * (1.0 + 0.05 * torch.randn(v.shape, generator=g))
                mk.load_state_dict(sdk)
                members.append(mk.eval())
            He, We = 2048 + 2 * 1792, W                     # three tile-rows of the Rwanda-shaped raster
            res_e = {}
            xe = synth_raster_slab(He, We, 777, dev)
            for tag, mods, share in (("one_member", members[:1], True), ("five_members", members, True), ("five_members_unshared", members, False)):
                e5 = ct.CountryEngine(mods, He, We, merge=not args.no_merge, rows_per_strip=args.rows_per_strip)
                if not share:
                    e5._bext_shared = False
                with torch.no_grad():
                    e5.run(xe, None, 1)
                    torch.cuda.synchronize()
                    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a0.record()
                    for _ in range(2):
                        e5.run(xe, None, 1)
                    a1.record()
                    torch.cuda.synchronize()
                res_e[tag] = a0.elapsed_time(a1) / 2
                res_e[tag + "_shared_builtup"] = bool(e5._bext_shared)
                del e5
            ens = {"raster": f"{He}x{We}", "ms_one_member": res_e["one_member"], "ms_five_members": res_e["five_members"],
                   "ms_five_members_unshared_builtup": res_e["five_members_unshared"],
                   "builtup_shared": res_e["five_members_shared_builtup"],
                   "value": He * We / (res_e["five_members"] * 1e-3), "unit": "pixels/s (5-member ensemble mean + std maps)",
                   "saving_from_sharing": 1.0 - res_e["five_members"] / res_e["five_members_unshared"],
                   "note": "N2: building_extractor never trains, so the ensemble's members carry identical copies -> one builtup pass per window"}
            del xe, members
            torch.cuda.empty_cache()
        except Exception as ex:
            ens = {"error": repr(ex)[:300]}

    # ---- BASELINE configs[4]: multi-temporal inference, 4 seasonal frames of a Switzerland-shaped raster, frames x row shards ----
    tseries = None
    if not args.skip_timeseries:
        try:
            from popcorn_b200 import timeseries as tsm
            Hs, Ws, T = 13408, 30592, 4
            torch.cuda.empty_cache()
            tse = tsm.TimeSeriesEngine([model], Hs, Ws, rank=rank, world=world, merge=True, rows_per_strip=args.rows_per_strip, frames=T)
            j0, j1 = tse.in_rows
            frame = synth_raster_slab(j1 - j0, Ws, j0, dev, seed=99) if j1 > j0 else torch.empty(6, 0, Ws, device=dev)
            frames = {t_: frame for t_ in tse.my_frames}    # same cost as independent draws; keeps 30 GB of generation out
            runs = []
            with torch.no_grad():
                tse.run(frames, None, 0, row_offset=j0)
                for _ in range(3):
                    barrier()
                    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a0.record()
                    o = tse.run(frames, None, 0, row_offset=j0)
                    a1.record()
                    barrier()
                    tms = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=dev)
                    if world > 1:
                        dist.all_reduce(tms, op=dist.ReduceOp.MAX)       # a run takes as long as its slowest rank
                    runs.append(float(tms.item()))
            tms = torch.tensor([sorted(runs)[1]], dtype=torch.float64, device=dev)      # median of three whole-series runs
            tseries = {"workload": f"switzerland_shaped_{Hs}x{Ws}_x{T}_seasonal_frames_over_{world}_gpus",
                       "partition": tse.describe(), "ms": float(tms.item()), "value": T * Hs * Ws / (float(tms.item()) * 1e-3), "unit": UNIT,
                       "season_total": float(o["season_total"].item()), "frames": T, "runs_ms": runs,
                       "note": "per-frame tiled inference + season mean/std/totals on the device (popcorn_b200.timeseries); inputs resident in HBM"}
            del frame, frames, o, tse
            torch.cuda.empty_cache()
        except Exception as ex:
            tseries = {"error": repr(ex)[:300]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": name, "H": H, "W": W, "regions": R_REGIONS, "patch": 2048, "overlap": 128,
                           "windows": "merged_row_strips" if not args.no_merge else "reference_tiles",
                           "rows_per_strip": args.rows_per_strip, "sharding": "balanced_256_row_units" if (balance and world > 1) else "strips", "ensemble": 1, "head": "dense",
                           "l2": "inputs (>=6 GB per GPU) far larger than the 126 MB L2; no flush needed",
                           "weights": weights_name,
                           "arithmetic": ("fp32 results: 3x3 convs and head as tensor-core products of "
                                          + ("fp16" if pb.weights.tc_operand_format() == 1 else "tf32")
                                          + " hi/lo operand halves (3 products per MAC, 22 significand bits, fp32 accumulation in TMEM); "
                                            "first conv layer, ConvT, epilogues, census sums (fp64) on the CUDA cores")},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "kernels": kernels,
                "fp32_simt": fp32,
                "cpu_baseline": cpu_base, "gpu_baseline": gpu_base, "train_step": train, "ensemble5": ens, "time_series": tseries,
                "same_slab_one_rank": alone, "check": check}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
