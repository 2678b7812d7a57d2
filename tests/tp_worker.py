"""Worker of tests/test_gpu_tp.py::test_tp_ipc_two_processes -- launched by torch.distributed.run, one process per
GPU.  Every rank loads its row shard, the ranks exchange CUDA IPC handles over a gloo group, then all evaluate the same
token stream in lock-step; rank 0 checks the logits against the oracle (CPU) bit for bit."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import llama_swift_b200 as lsb  # noqa: E402
from llama_swift_b200 import dist_util  # noqa: E402


def main():
    path, n_ctx = sys.argv[1], int(sys.argv[2])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    model = dist_util.tp_load(lsb, path, n_ctx, int(os.environ.get("LOCAL_RANK", rank)))
    rng = np.random.default_rng(7)
    n_past, got, calls = 0, [], []
    for n in (4, 9, 1, 1, 1, 1):
        toks = rng.integers(3, 512, size=n).astype(np.int32)
        got.append(lsb.llama_eval(model, 8, n_past, toks))
        calls.append((n_past, toks))
        n_past += n
    stream = rng.integers(3, 512, size=9).astype(np.int32)
    dtoks, dlogits, _ = model.decode_device(n_past, int(stream[0]), 8, n_threads=8, forced_tokens=stream[1:], want_logits=True)
    dist.barrier()
    # all ranks hold identical full logits
    mine = np.stack(got)
    ref = [None] * world
    dist.all_gather_object(ref, mine.tobytes())
    assert all(r == ref[0] for r in ref), "ranks disagree on the logits"
    if rank == 0:
        import ctypes as C
        from conftest import CpuModel, _bind_model_api, ORACLE_SO, bits, rel_l2
        L = C.CDLL(ORACLE_SO)
        _bind_model_api(L, "ora")
        ora = CpuModel(L, "ora", path, n_ctx)
        exact = 0
        for (p, toks), g in zip(calls, got):
            want = ora.eval(8, p, toks)
            assert rel_l2(g, want) <= 1e-3 and int(g.argmax()) == int(want.argmax()), (p, rel_l2(g, want))
            exact += int(np.array_equal(bits(g), bits(want)))
        for i in range(8):
            want = ora.eval(8, n_past + i, stream[i:i + 1])
            assert rel_l2(dlogits[i], want) <= 1e-3 and int(dtoks[i]) == int(want.argmax())
            exact += int(np.array_equal(bits(dlogits[i]), bits(want)))
        print(f"TP_WORKER world={world}: {exact}/{len(calls) + 8} evaluations bit-identical to the oracle", flush=True)
        assert exact >= len(calls) + 8 - 1
        ora.free()
    dist.barrier()
    model.free()
    print(f"TP_WORKER_OK rank {rank}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
