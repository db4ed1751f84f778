"""CPU: the pieces of bench.py's JSON contract that are pure host code — per-kernel byte / FLOP models, the roofline object, the ncu
traffic lookup and the kernel-name mapping of tools/ncu_traffic.py (no GPU, no oracle)."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
import ncu_traffic  # noqa: E402


def test_conv_work_models():
    """(Cin + Cout) * 4 B and 2 * 9 * Cin * Cout FLOP per pixel; pooled / dot / ConvT epilogues add their own outputs."""
    assert bench._conv_work("conv3x3_tc<8,0,8,store>") == (2 * 9 * 8 * 8, 64)
    assert bench._conv_work("conv3x3_tc<8,8,8,store>") == (2 * 9 * 16 * 8, 96)
    assert bench._conv_work("conv3x3_tc<16,16,8,store>") == (2 * 9 * 32 * 8, 160)
    assert bench._conv_work("conv3x3_tc<8,0,8,pool>") == (2 * 9 * 8 * 8, 32 + 32 + 8)
    assert bench._conv_work("conv3x3_tc<8,0,8,dot>") == (2 * 9 * 8 * 8, 32 + 8)
    f, b = bench._conv_work("conv3x3_tc<16,0,16,convt>")
    assert f == 2 * 9 * 16 * 16 + 2 * 4 * 16 * 16 and b == 64 + 4 * 16 * 4
    assert bench._conv_work("conv3x3<2,0,8,store>") == (2 * 9 * 2 * 8, 8 + 32)
    assert bench._conv_work("conv3x3_tc<4,0,8,store>") == (2 * 9 * 4 * 8, 16 + 32)


def test_kernel_table_and_roofline():
    """Every row carries its own bound and fraction; the roofline object is the row with the largest time, with the measured peak it
    was divided by and the ncu-derived traffic scaled to one launch."""
    px = 64.0e6
    prof = {
        "conv3x3_tc<8,8,8,store>": (20.0, 20, 20 * px),        # (ms, launches, pixels)
        "head_tc<dense>": (18.0, 10, 10 * px),
        "conv3x3<4,0,8,store>": (5.0, 10, 10 * px / 2),
        "accumulate": (2.0, 10, 10 * px),
        "region_sum": (0.4, 1, px * 4),
        "never_ran": (0.0, 0, 0.0),
    }
    rows = bench.kernel_table(prof, ms_total=50.0, hbm_peak=6500.0, tensor_peak=1400.0, fp32_peak=70.0, frac_multi=0.05)
    names = [r["kernel"] for r in rows]
    assert names[0] == "conv3x3_tc<8,8,8,store>" and "never_ran" not in names and names == sorted(names, key=lambda n: -prof[n][0])
    by = {r["kernel"]: r for r in rows}
    c = by["conv3x3_tc<8,8,8,store>"]
    assert c["bound"] == "hbm" and c["unit"] == "GB/s" and c["peak"] == 6500.0
    assert c["achieved"] == pytest.approx(96 * 20 * px / 20.0e-3 / 1e9) and c["frac"] == pytest.approx(c["achieved"] / 6500.0)
    assert c["algorithmic_bytes"] == pytest.approx(96 * px) and c["share_of_step"] == pytest.approx(0.4)
    h = by["head_tc<dense>"]
    assert h["bound"] == "tensor" and h["unit"] == "TFLOP/s" and h["peak"] == 1400.0
    assert h["achieved"] == pytest.approx(bench.FLOP_PER_PX_HEAD * 10 * px / 18.0e-3 / 1e12)
    assert h["frac_of_split3_ceiling"] == pytest.approx(h["achieved"] / h["ceiling_split3"]) and h["operands"] in ("fp16 hi/lo", "tf32 hi/lo")
    assert by["conv3x3<4,0,8,store>"]["bound"] == "hbm" and by["accumulate"]["bound"] == "hbm"
    traffic = bench.ncu_traffic_per_px()
    if "conv3x3_tc<8,8,8,store>" in traffic:      # the committed capture: DRAM bytes per pixel ~ algorithmic bytes (no re-reads)
        assert c["traffic"] == pytest.approx(traffic["conv3x3_tc<8,8,8,store>"] * px)
        assert 0.9 < c["traffic"] / c["algorithmic_bytes"] < 1.1
    r = bench.roofline_of(rows, "measured")
    assert r["kernel"] == "conv3x3_tc<8,8,8,store>" and r["bound"] == "hbm" and r["frac"] == c["frac"] and r["peak_source"] == "measured"
    assert r["launches"] == 20 and r["avg_launch_ms"] == pytest.approx(1.0) and r["traffic"] == c["traffic"]
    json.dumps(rows), json.dumps(r)                # the line must serialise


def test_ncu_kernel_name_mapping():
    """Demangled kernel names of an ncu CSV -> the names of bench.py's kernel table (the join key of roofline.traffic)."""
    m = ncu_traffic.bench_name
    assert m("void pc::conv3x3_tc_kernel<8, 8, 8, 0>(TcConvParams)") == "conv3x3_tc<8,8,8,store>"
    assert m("void conv3x3_tc_kernel<16, 0, 16, 3>(TcConvParams)") == "conv3x3_tc<16,0,16,convt>"
    assert m("void conv3x3_tc_kernel<2, 0, 8, 0>(TcConvParams)") == "conv3x3_tc<2,0,8,store>"
    assert m("void conv3x3_kernel<4, 0, 8, 0, 1, 0>(ConvParams)") == "conv3x3<4,0,8,store>"
    assert m("void head_tc_kernel<16, 0, 1>(HeadArgs)") == "head_tc<dense>"
    assert m("void head_tc_kernel<(int)16, (bool)1, (bool)1>(HeadArgs)") == "head_tc<sparse>"
    assert m("void accumulate_kernel<1>(const float *, const float *, int)") == "accumulate"
    assert m("void finalize_kernel<true>(float *, float *)") == "finalize"
    assert m("void region_sum_kernel<1>(const float *, const int *, long long, int, double *)") == "region_sum"
    assert m("void at::native::vectorized_elementwise_kernel<4, FillFunctor<float>>(int)") is None


def test_committed_traffic_capture_matches_the_byte_models():
    """profiles/r2_ncu_bench_traffic.json: per kernel, measured DRAM bytes per pixel within 10 % of the algorithmic model (nothing is
    re-read inside a kernel) — what DESIGN.md §5 states."""
    traffic = bench.ncu_traffic_per_px()
    if not traffic:
        pytest.skip("no committed ncu capture")
    for name, bpp in traffic.items():
        if name.startswith("conv3x3"):
            _, model = bench._conv_work(name)
            assert 0.9 < bpp / model < 1.1, (name, bpp, model)
    assert 0.95 < traffic["head_tc<dense>"] / bench.BYTES_PER_PX_HEAD < 1.05
