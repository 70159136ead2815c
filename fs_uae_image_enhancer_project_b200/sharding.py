"""Frame-wise sharding of a framebuffer stream over the GPUs of one box.

Frames are independent -- no halo crosses frames, nothing is reduced -- so N GPUs are N engine
replicas, each fed a contiguous frame range; there is deliberately no data-path collective and NCCL
is not used for data.  ``torch.distributed`` (nccl on GPUs, gloo in the CPU tests) only provides the
barrier and the max-over-ranks of the measured time.
"""
from __future__ import annotations

from typing import Optional, Tuple

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
