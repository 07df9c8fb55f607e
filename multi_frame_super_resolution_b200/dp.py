"""Data-parallel host logic: independent bursts shard across ranks with NO data-path collective.

Used by bench.py (--mode dp) and covered under gloo by tests/test_dp_cpu.py; the row-band split of ONE large burst lives in rowband.py.
The reference is single-GPU (cudaSetDevice(0), test_opencv/kernel.cu:45).  Bursts are independent
units, so burst b simply goes to rank b mod G; the only communication is the
aggregation of the timing/throughput scalars, which works on any torch.distributed backend
(NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def shard_bursts(n_bursts: int, rank: int, world: int) -> List[int]:
    """Burst ids owned by `rank`: round-robin, so every rank gets floor or ceil of n/world."""
    if world < 1 or not (0 <= rank < world) or n_bursts < 0:
        raise ValueError("bad shard arguments")
    return list(range(rank, n_bursts, world))


def burst_seed(base_seed: int, burst_id: int) -> int:
    """burst b uses seed base + b (so any rank generates the same burst b)."""
    return base_seed + burst_id


def aggregate_throughput(local_units: float, local_ms: float, device="cpu"):
    """Whole-job throughput: units processed by all ranks / max-over-ranks time.

    Returns (units_total, ms_max, units_per_second).  With no process group it is the local value."""
    u = torch.tensor([float(local_units)], dtype=torch.float64, device=device)
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return float(u.item()), ms, (float(u.item()) / (ms / 1e3) if ms > 0 else float("inf"))
