"""DRAM traffic of bench.py's kernels from an ncu capture of the SAME command (development / evidence tool).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/ncu_bench.csv python bench.py --steps 1 --warmup 3 --skip-e2e --skip-train ... > gpurun_out/ncu_bench.log
    python tools/ncu_traffic.py gpurun_out/ncu_bench.csv gpurun_out/ncu_bench.log profiles/r2_ncu_bench_traffic.json

Per kernel (named as in bench.py's `kernels` table): launches, total DRAM bytes read + written over all captured launches, the share
of the summed kernel time, and — with the pixels each launch processed taken from the bench line of the same run —
`dram_bytes_per_px`, which bench.py multiplies by the pixels of a launch to fill `roofline.traffic`.  ncu serialises the launches and
replays them cold, so durations are only good for SHARES; bytes are exact.
"""
import csv
import json
import re
import sys

EPI = {"0": "store", "1": "pool", "2": "dot", "3": "convt"}


def bench_name(kernel: str):
    k = kernel.replace(" ", "")
    m = re.search(r"conv3x3_tc_kernel<(\d+),(\d+),(\d+),(\d+)>", k)
    if m:
        return f"conv3x3_tc<{m[1]},{m[2]},{m[3]},{EPI[m[4]]}>"
    m = re.search(r"conv3x3_kernel<(\d+),(\d+),(\d+),(\d+)", k)
    if m:
        return f"conv3x3<{m[1]},{m[2]},{m[3]},{EPI[m[4]]}>"
    m = re.search(r"convt2x2_kernel<(\d+)>", k)
    if m:
        return f"convt2x2<{m[1]}>"
    m = re.search(r"head_tc_kernel<(?:\(int\))?(\d+),(?:\(bool\))?(\d+|true|false)[,>]", k)
    if m:
        return "head_tc<sparse>" if m[2] in ("1", "true") else "head_tc<dense>"
    for pat, name in (("head_backward_kernel", "head_backward"), ("head_bwd_reduce", "head_backward"), ("accumulate_kernel", "accumulate"),
                      ("finalize_kernel", "finalize"), ("region_sum_kernel", "region_sum"), ("ingest_kernel", "ingest_normalize"),
                      ("compact_", "sparse_mask_compact"), ("head_forward_kernel", "head_forward_simt<dense>")):
        if pat in k:
            return name
    return None


def parse(csv_path):
    rows = []
    with open(csv_path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    per = {}
    for r in rd:
        name = r.get("Kernel Name") or r.get("Kernel")
        metric = r.get("Metric Name")
        if not name or not metric:
            continue
        val = float(str(r.get("Metric Value", "0")).replace(",", "") or 0)
        unit = (r.get("Metric Unit") or "").lower()
        key = r.get("ID")
        d = per.setdefault(key, {"kernel": name, "read": 0.0, "write": 0.0, "ns": 0.0})
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, None)
        if metric == "dram__bytes_read.sum":
            d["read"] = val * (scale or 1.0)
        elif metric == "dram__bytes_write.sum":
            d["write"] = val * (scale or 1.0)
        elif metric == "gpu__time_duration.sum":
            d["ns"] = val * {"nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1.0)
    rows = list(per.values())
    return rows


def main():
    csv_path, log_path, out_path = sys.argv[1:4]
    rows = parse(csv_path)
    line = None
    for ln in open(log_path):
        if ln.startswith("{") and '"kernels"' in ln:
            line = json.loads(ln)
    if line is None:
        raise SystemExit("no bench JSON line in " + log_path)
    steps_total = line["steps"] + line["warmup"]
    px_per_step = {k["kernel"]: k["pixels"] / line["steps"] for k in line["kernels"]}
    agg = {}
    other = {"launches": 0, "ns": 0.0, "bytes": 0.0}
    for r in rows:
        n = bench_name(r["kernel"])
        tgt = agg.setdefault(n, {"launches": 0, "ns": 0.0, "read": 0.0, "write": 0.0}) if n else None
        if tgt is None:
            other["launches"] += 1; other["ns"] += r["ns"]; other["bytes"] += r["read"] + r["write"]
            continue
        tgt["launches"] += 1; tgt["ns"] += r["ns"]; tgt["read"] += r["read"]; tgt["write"] += r["write"]
    tot_ns = sum(v["ns"] for v in agg.values()) + other["ns"]
    uniq = line["config"]["H"] * line["config"]["W"] / max(line["n_gpus"], 1)
    out = {"command": "bench.py " + " ".join(sys.argv[4:]), "steps_captured": steps_total, "unique_px_per_step": uniq, "kernels": {},
           "other_launches": other}
    path_bytes = 0.0
    for n, v in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        px = px_per_step.get(n)
        per_step_bytes = (v["read"] + v["write"]) / steps_total
        path_bytes += per_step_bytes
        out["kernels"][n] = {"launches_per_step": v["launches"] / steps_total, "dram_read_bytes_per_step": v["read"] / steps_total,
                             "dram_write_bytes_per_step": v["write"] / steps_total, "share_of_kernel_time": v["ns"] / tot_ns if tot_ns else None,
                             "dram_bytes_per_px": (per_step_bytes / px) if px else None, "pixels_per_step": px}
    out["path_dram_bytes_per_unique_px"] = path_bytes / uniq
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps({"path_dram_bytes_per_unique_px": out["path_dram_bytes_per_unique_px"], "kernels": len(out["kernels"]),
                      "unmatched_launches": other["launches"]}))
    for n, v in out["kernels"].items():
        print(f"{n:30s} launches/step {v['launches_per_step']:6.1f}  share {100 * (v['share_of_kernel_time'] or 0):5.1f} %  "
              f"dram B/px {v['dram_bytes_per_px'] if v['dram_bytes_per_px'] is None else round(v['dram_bytes_per_px'], 2)}")


if __name__ == "__main__":
    main()
