"""Tensor-parallel groups on 2+ GPUs: rows of every matrix split over the group, activation
slices all-gathered by peer stores inside the token kernel.  Because a row is a complete reference dot product, the
group's logits must equal the single-GPU logits -- and the oracle's -- bit for bit.  Skipped on a 1-GPU box."""
import os
import subprocess
import sys

import numpy as np
import pytest

import llama_swift_b200 as lsb
from conftest import ROOT, CpuModel, bits, rel_l2

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("tp", [2, 4, 8])
def test_tp_group_single_process(oracle_lib, small_model, tp):
    """b200_llama_load_group: one host thread drives all GPUs of the group (what a LlamaRunner app would do)."""
    if _n_gpus() < tp:
        pytest.skip(f"needs {tp} GPUs")
    ora = CpuModel(oracle_lib, "ora", small_model, 64)
    one = lsb.llama_model_load(small_model, n_ctx=64, device=0)
    grp = lsb.llama_model_load_group(small_model, n_ctx=64, devices=tuple(range(tp)))
    try:
        rng = np.random.default_rng(11)
        n_past, exact = 0, 0
        steps = (4, 9, 1, 1, 1, 1)
        for n in steps:
            toks = rng.integers(3, 512, size=n).astype(np.int32)
            want = ora.eval(8, n_past, toks)
            a = lsb.llama_eval(one, 8, n_past, toks)
            g = lsb.llama_eval(grp, 8, n_past, toks)
            assert np.array_equal(bits(g), bits(a)), f"tp={tp} differs from 1 GPU at n_past={n_past}: rel {rel_l2(g, a):.3e}"
            assert rel_l2(g, want) <= 1e-3 and g.argmax() == want.argmax()
            exact += int(np.array_equal(bits(g), bits(want)))
            n_past += n
        assert exact >= len(steps) - 1
        stream = rng.integers(3, 512, size=13).astype(np.int32)
        t1, l1, _ = one.decode_device(n_past, int(stream[0]), 12, n_threads=8, forced_tokens=stream[1:], want_logits=True)
        t2, l2, ms = grp.decode_device(n_past, int(stream[0]), 12, n_threads=8, forced_tokens=stream[1:], want_logits=True)
        assert np.array_equal(t1, t2) and np.array_equal(bits(l1), bits(l2))
        print(f"[tp] group of {tp}: bit-identical to 1 GPU; 12 device-resident steps in {ms:.3f} ms")
    finally:
        ora.free()
        one.free()
        grp.free()


def test_tp_group_long_enqueue(small_model):
    """One host thread, many steps: 600 device-resident steps and a 600-token llama_eval on a single-process group of 2.  Every
    rank's token kernel waits for its peers ON THE GPU, so the host must hand token i to every rank before it queues enough
    work on one rank to block in a launch (the driver's pending-launch queue holds ~1K entries) -- the enqueue is token-major
    with a synchronisation every 64 tokens (engine.cu).  Results must equal the single GPU's."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    n_ctx = 640
    one = lsb.llama_model_load(small_model, n_ctx=n_ctx, device=0)
    grp = lsb.llama_model_load_group(small_model, n_ctx=n_ctx, devices=(0, 1))
    try:
        rng = np.random.default_rng(17)
        stream = rng.integers(3, 512, size=601).astype(np.int32)
        t1, _, _ = one.decode_device(0, int(stream[0]), 600, n_threads=8, forced_tokens=stream[1:])
        t2, _, ms = grp.decode_device(0, int(stream[0]), 600, n_threads=8, forced_tokens=stream[1:])
        assert np.array_equal(t1, t2)
        one.set_option("batch", 0)                  # the group evaluates token by token; compare like with like
        a = lsb.llama_eval(one, 8, 0, stream[:600])
        g = lsb.llama_eval(grp, 8, 0, stream[:600])
        assert np.array_equal(bits(g), bits(a))
        print(f"[tp] group of 2, single process: 600 device-resident steps ({ms:.1f} ms) and a 600-token llama_eval identical to 1 GPU")
    finally:
        one.free()
        grp.free()


@pytest.mark.parametrize("tp", [2, 4, 8])
def test_tp_group_13b_shapes(oracle_lib, tp):
    """BASELINE.json configs[3]: LLaMA-13B geometry (two-part file, 40 heads, n_ff 13824) row-sharded over the group."""
    if _n_gpus() < tp:
        pytest.skip(f"needs {tp} GPUs")
    from conftest import model_file
    path = model_file(n_layer=1, n_vocab=512, seed=31, n_embd=5120)
    ora = CpuModel(oracle_lib, "ora", path, 32)
    grp = lsb.llama_model_load_group(path, n_ctx=32, devices=tuple(range(tp)))
    try:
        rng = np.random.default_rng(13)
        n_past, exact = 0, 0
        steps = (4, 9, 1, 1, 1)
        for n in steps:
            toks = rng.integers(3, 512, size=n).astype(np.int32)
            want = ora.eval(8, n_past, toks)
            g = lsb.llama_eval(grp, 8, n_past, toks)
            assert rel_l2(g, want) <= 1e-3 and g.argmax() == want.argmax()
            exact += int(np.array_equal(bits(g), bits(want)))
            n_past += n
        print(f"[tp] 13B shapes over {tp} GPUs: {exact}/{len(steps)} evaluations bit-identical to the oracle")
        assert exact >= len(steps) - 1
    finally:
        ora.free()
        grp.free()


def test_tp_ipc_two_processes(oracle_lib, small_model):
    """One process per GPU (the bench.py / torchrun launch): IPC handles over gloo, lock-step evaluation."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "tp_worker.py"), small_model, "64"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0
    assert "TP_WORKER_OK rank 0" in r.stdout and "TP_WORKER_OK rank 1" in r.stdout


def test_tp_needs_connect(small_model):
    """A shard that was never connected refuses to evaluate (no silent single-GPU fallback)."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    m = lsb.llama_model_load_shard(small_model, 64, 0, 0, 2)
    try:
        with pytest.raises(lsb.LlamaError):
            lsb.llama_eval(m, 8, 0, np.array([1, 2], np.int32))
    finally:
        m.free()
