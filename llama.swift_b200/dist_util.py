"""Multi-rank plumbing (torch.distributed is plumbing here, not the product).

The decode path shards as tensor parallelism with every matrix split by ROWS over the GPUs of one NVSwitch box
(DESIGN.md section 6): the data-path exchange -- an all-gather of finished activation slices --
is done by peer-to-peer stores from inside the token kernel, so the only things torch.distributed carries are the
one-off exchange of the 64-byte CUDA IPC handles at load time, barriers, and the timing reduction (max over ranks)."""
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


def gather_handles(mine: bytes) -> list:
    """Every rank's opaque handle, in rank order (one all_gather_object; works on gloo and nccl groups)."""
    if not (dist.is_available() and dist.is_initialized()):
        return [mine]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, bytes(mine))
    return out


def tp_load(lsb, path: str, n_ctx: int, device: int):
    """One-process-per-GPU tensor-parallel load: this rank's row shard on `device`, IPC handles exchanged over the
    default process group, peers mapped, barrier.  Returns a LlamaModel whose llama_eval / decode_device must then be
    called with identical arguments on every rank."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return lsb.llama_model_load(path, n_ctx=n_ctx, device=device)
    model = lsb.llama_model_load_shard(path, n_ctx, device, rank, world)
    handles = gather_handles(lsb.tp_ipc_handle(model))
    lsb.tp_connect(model, handles)
    dist.barrier()
    return model
