"""Multi-rank plumbing for the benchmark harness (torch.distributed is plumbing here, not the product).

The decode path shards as Megatron-style tensor parallelism; until that is built the ranks of a
`torchrun` launch are independent replicas, and the only cross-rank step is the timing reduction: barrier,
max-over-ranks of the device-timed region, units summed over ranks."""
from __future__ import annotations

import torch
import torch.distributed as dist


def barrier_and_sync(device_sync) -> None:
    """barrier + device synchronize on both sides of a timed region (device_sync: callable, e.g. torch.cuda.synchronize)."""
    device_sync()
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    device_sync()


def max_over_ranks(value: float, device: str = "cpu") -> float:
    """Device time of a timed region, taken as the maximum over all ranks."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: str = "cpu") -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(units_this_rank: float, seconds_this_rank: float, device: str = "cpu") -> float:
    """Whole-job throughput: units processed by all ranks / max-over-ranks time."""
    return sum_over_ranks(units_this_rank, device) / max_over_ranks(seconds_this_rank, device)
