"""CPU: the round-2 candidate kernels under tools/probe/ are not on the shipped path and have never run on hardware; what CAN be
checked without a GPU is kept green here — their index math (numpy emulation of packing, staged rows, UMMA windows and the accumulator
ring, exact against conv2d) and their warp-role protocol (randomised mbarrier simulation: no deadlock, stage / slot invariants)."""
import os
import subprocess
import sys

import pytest

PROBE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "probe")


@pytest.mark.parametrize("args", [["emulate_conv_pair.py"], ["simulate_protocol.py"], ["simulate_protocol.py", "--pair"]])
def test_probe_tool_runs_clean(args):
    r = subprocess.run([sys.executable, os.path.join(PROBE, args[0])] + args[1:], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    if args[0].startswith("emulate"):
        errs = [float(line.split("max err")[1].split()[-1] if "vs 3xTF32" not in line else line.split("formula")[1].split()[0])
                for line in r.stdout.splitlines() if "max err" in line]
        assert errs and max(errs) < 1e-12, r.stdout
    else:
        assert "no deadlock" in r.stdout
