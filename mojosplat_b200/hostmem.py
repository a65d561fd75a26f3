"""Host-side placement for the end-to-end (host buffers in / host image out) path on a multi-GPU box.

The per-frame upload (56 B per Gaussian) and download (12 B per pixel) run at PCIe speed only when the pinned host
buffers live on the NUMA node the GPU hangs off and the copying process runs there too; a box whose ranks all sit on
node 0 funnels every transfer of the remote GPUs through the inter-socket links.  ``bind_to_device_node`` is best
effort: containers often restrict the CPU set or the memory nodes, in which case it reports what it could not do.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import torch

MPOL_PREFERRED = 1
_SYS_set_mempolicy = 238  # x86_64


def device_pci_address(index: int) -> str:
    p = torch.cuda.get_device_properties(index)
    return f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"


def device_numa_node(index: int) -> int:
    """NUMA node of GPU ``index`` (-1 when the platform does not say)."""
    try:
        return int(Path(f"/sys/bus/pci/devices/{device_pci_address(index)}/numa_node").read_text().strip())
    except Exception:
        return -1


def node_cpus(node: int) -> set:
    cpus = set()
    try:
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
    except Exception:
        pass
    return cpus


def bind_to_device_node(index: int) -> dict:
    """Run this process, and place the memory it allocates from now on (pinned buffers included), on the NUMA node
    of GPU ``index``.  Returns what was achieved: {"node", "cpus_bound", "mem_bound", "note"}."""
    node = device_numa_node(index)
    out = {"node": node, "cpus_bound": False, "mem_bound": False, "note": ""}
    if node < 0:
        out["note"] = "GPU NUMA node unknown"
        return out
    want = node_cpus(node)
    try:
        allowed = os.sched_getaffinity(0)
        pick = want & allowed
        if pick:
            os.sched_setaffinity(0, pick)
            out["cpus_bound"] = True
        else:
            out["note"] += f"no allowed CPU on node {node} (allowed: {min(allowed)}-{max(allowed)}); "
    except Exception as e:  # pragma: no cover
        out["note"] += f"sched_setaffinity: {e}; "
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(_SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(64))
        if rc == 0:
            out["mem_bound"] = True
        else:
            out["note"] += f"set_mempolicy errno {ctypes.get_errno()}; "
    except Exception as e:  # pragma: no cover
        out["note"] += f"set_mempolicy: {e}; "
    return out
