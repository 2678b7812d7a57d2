"""World-size-2 gloo test of the N>1 harness logic (runs on CPU): max-over-ranks timing, whole-job aggregation."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from llama_swift_b200 import dist_util
    dist_util.barrier_and_sync(lambda: None)
    secs = 0.5 if rank == 0 else 0.8            # the slower rank defines the job time
    t = dist_util.max_over_ranks(secs)
    thr = dist_util.aggregate_throughput(512.0, secs)
    out.put((rank, t, thr))
    dist.destroy_process_group()


def test_two_rank_aggregation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, t, thr in res:
        assert abs(t - 0.8) < 1e-12
        assert abs(thr - 1024.0 / 0.8) < 1e-9


def _handles_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from llama_swift_b200 import dist_util
    mine = bytes([rank]) * 64                    # stands in for a 64-byte cudaIpcMemHandle_t
    got = dist_util.gather_handles(mine)
    out.put((rank, [h[0] for h in got], [len(h) for h in got]))
    dist.destroy_process_group()


def test_tp_handle_exchange_rank_order():
    """The one thing torch.distributed carries for a tensor-parallel group: every rank's IPC handle, in rank order."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_handles_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, firsts, lens in res:
        assert firsts == [0, 1] and lens == [64, 64]
