"""Host-side placement for the host-buffer (e2e) path on multi-GPU boxes: bind the calling process to the CPU
cores of the NUMA node its GPU hangs off BEFORE it allocates pinned buffers, so that first-touch puts the pinned
pages in that node's memory and the PCIe copies do not cross the socket interconnect."""
from __future__ import annotations

import os
from typing import List, Optional


def _cpulist(text: str) -> List[int]:
    out: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            out.extend(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out


def gpu_numa_node(device_index: int) -> Optional[int]:
    """NUMA node of a CUDA device (from its PCI address), or None when the platform does not say."""
    try:
        import torch

        p = torch.cuda.get_device_properties(device_index)
        addr = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{addr}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_node(device_index: int) -> dict:
    """Restrict this process to the cores of the GPU's NUMA node.  Returns what was done (for logs)."""
    node = gpu_numa_node(device_index)
    info = {"numa_node": node, "bound": False}
    if node is None:
        return info
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _cpulist(f.read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cores=len(allowed))
    except Exception as ex:  # containers may forbid it: the run goes on unbound
        info["error"] = repr(ex)
    return info
