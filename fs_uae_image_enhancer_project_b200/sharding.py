"""Frame-wise sharding of a framebuffer stream over the GPUs of one box.

Frames are independent -- no halo crosses frames, nothing is reduced -- so N GPUs are N engine
replicas, each fed a contiguous frame range; there is deliberately no data-path collective and NCCL
is not used for data.  ``torch.distributed`` (nccl on GPUs, gloo in the CPU tests) only provides the
barrier and the max-over-ranks of the measured time.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def frame_range(n_frames: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous range [start, stop) of rank ``rank``: GPU g of G gets frames [g*N/G, (g+1)*N/G)."""
    if world < 1 or not 0 <= rank < world or n_frames < 0:
        raise ValueError("bad shard request")
    return (n_frames * rank) // world, (n_frames * (rank + 1)) // world


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    """Whole-job time = the slowest rank's device time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_sharded(model, frames_host: torch.Tensor, out_host: torch.Tensor, rank: int, world: int, device: int,
                gamma: bool = True, crop16: bool = False) -> Tuple[int, int]:
    """Process this rank's slice of a host-resident uint8 [N,H,W,4] stream in place of ``out_host``."""
    lo, hi = frame_range(frames_host.shape[0], world, rank)
    if hi > lo:
        model.run_host(frames_host[lo:hi], out_host[lo:hi], gamma=gamma, crop16=crop16, device=device)
    return lo, hi


def gpu_local_cpus(device: int) -> List[int]:
    """CPUs of the NUMA node the GPU's PCIe root hangs off (sysfs), [] when the platform does not say."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device)).busId
        bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()
        if len(bdf.split(":")[0]) == 8:                      # NVML prints an 8-digit PCI domain, sysfs uses 4
            bdf = bdf[4:]
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as fh:
            text = fh.read().strip()
        cpus: List[int] = []
        for part in text.split(","):
            if part:
                lo, _, hi = part.partition("-")
                cpus.extend(range(int(lo), int(hi or lo) + 1))
        return cpus
    except Exception:
        return []


def bind_to_gpu_numa_node(device: int) -> List[int]:
    """Pin this process to the CPUs next to its GPU *before* it allocates pinned host buffers, so first-touch places
    them on the GPU's NUMA node: with one process per GPU, host staging then never crosses the socket interconnect.
    Best effort; returns the CPU list used ([] = nothing changed)."""
    cpus = gpu_local_cpus(device)
    allowed = sorted(set(cpus) & set(os.sched_getaffinity(0))) if cpus else []
    if allowed and len(allowed) < len(os.sched_getaffinity(0)):
        os.sched_setaffinity(0, allowed)
        return allowed
    return []
