"""Bind a rank's host threads (and therefore its first-touch pinned allocations) to the NUMA node of its GPU.

On an 8-GPU box the GPUs hang off two CPU sockets; pinned staging buffers that all live on node 0 make the ranks on the other
socket pull their host->device copies across the inter-socket link (round 1: 13 GB/s per GPU at N = 8 against 22.5 GB/s at
N = 1).  The reference is single-GPU and has no equivalent; this is host plumbing of the country driver.
"""
from __future__ import annotations

import os
from typing import Optional


def _read(path: str) -> Optional[str]:
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def parse_cpulist(text: str):
    """'0-3,8,10-11' -> {0,1,2,3,8,10,11}"""
    cpus = set()
    for part in text.split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(device_index: int) -> Optional[int]:
    """NUMA node of a CUDA device from its PCI address in sysfs; None when the platform does not say (or says -1)."""
    import torch
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
    except Exception:
        return None
    node = _read(f"/sys/bus/pci/devices/{bus}/numa_node")
    if node is None:
        return None
    try:
        n = int(node)
    except ValueError:
        return None
    return n if n >= 0 else None


def bind_to_gpu_node(device_index: int) -> dict:
    """sched_setaffinity of this process to the CPUs of the GPU's NUMA node (intersected with the CPUs it may already use).
    Pages of later pinned allocations are then first-touched on that node.  Returns what was done (for the bench line)."""
    info = {"node": None, "cpus": None, "bound": False}
    node = gpu_numa_node(device_index)
    info["node"] = node
    if node is None:
        return info
    text = _read(f"/sys/devices/system/node/node{node}/cpulist")
    if not text:
        return info
    try:
        allowed = os.sched_getaffinity(0)
        want = parse_cpulist(text) & allowed
        if want and want != allowed:
            os.sched_setaffinity(0, want)
            info["bound"] = True
        info["cpus"] = len(want or allowed)
    except (OSError, AttributeError, ValueError):
        pass
    return info
