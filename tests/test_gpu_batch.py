"""Prompt batches (llama_eval with N > 1): the layer-by-layer batch path (csrc/batch.cuh: every weight row read once per
group of columns, like ggml.c:6199-6222) against the oracle and against the token-by-token path.  Bit-identical logits and
KV cache are demanded: a column of a batched mat-mul is the same sequence of operations as the single-token mat-vec."""
import numpy as np
import pytest

import llama_swift_b200 as lsb
from conftest import CpuModel, bits, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tc", [1, 0], ids=["tcgen05", "cuda-core-columns"])
@pytest.mark.parametrize("n_threads", [8, 3])
def test_batch_vs_oracle(oracle_lib, small_model, n_threads, tc):
    """N = 2 / 4 / 9 (the reference's own batch sizes, PO.mm:822, 885) / 17 / 64, interleaved with single tokens; the mat-mul
    on the tensor cores (tcgen05 + TMEM, csrc/prefill_tc.cuh) and on the CUDA-core multi-column loop."""
    ora = CpuModel(oracle_lib, "ora", small_model, 128)
    gpu = lsb.llama_model_load(small_model, n_ctx=128)
    gpu.set_option("tc", tc)
    gpu.set_option("tc_min_n", 2)          # every batch size through the tensor-core kernel (the default switches at 24 tokens)
    try:
        rng = np.random.default_rng(23)
        n_past, exact, total = 0, 0, 0
        for n in (4, 9, 1, 2, 17, 1, 64, 1):
            toks = rng.integers(3, 512, size=n).astype(np.int32)
            want = ora.eval(n_threads, n_past, toks)
            got = lsb.llama_eval(gpu, n_threads, n_past, toks)
            r = rel_l2(got, want)
            same = np.array_equal(bits(got), bits(want))
            print(f"[batch] tc={tc} nth={n_threads} n_past={n_past} N={n}: bit-identical={same} rel_l2={r:.3e} launches={gpu.last_launches}")
            assert r <= 1e-3 and int(got.argmax()) == int(want.argmax())
            exact += int(same)
            total += 1
            n_past += n
        assert exact >= total - 1
        for il in range(2):
            for which in (0, 1):
                g, w = gpu.kv_export(il, which, n_past), ora.kv(il, which, n_past)
                assert np.array_equal(bits(g), bits(w)) or rel_l2(g, w) <= 1e-6
    finally:
        ora.free()
        gpu.free()


def test_batch_equals_token_by_token(small_model):
    """300 tokens in one call (two internal chunks of the batch path) == the same call evaluated one token at a time."""
    n = 300
    a = lsb.llama_model_load(small_model, n_ctx=320)
    b = lsb.llama_model_load(small_model, n_ctx=320)
    try:
        b.set_option("batch", 0)
        rng = np.random.default_rng(5)
        toks = rng.integers(3, 512, size=n).astype(np.int32)
        la = lsb.llama_eval(a, 8, 0, toks)
        lb = lsb.llama_eval(b, 8, 0, toks)
        assert np.array_equal(bits(la), bits(lb))
        for il in range(2):
            for which in (0, 1):
                assert np.array_equal(bits(a.kv_export(il, which, n)), bits(b.kv_export(il, which, n)))
        # and a decode step on top of both caches
        t = np.array([int(la.argmax())], np.int32)
        assert np.array_equal(bits(lsb.llama_eval(a, 8, n, t)), bits(lsb.llama_eval(b, 8, n, t)))
        print(f"[batch] N={n}: batch path {a.last_launches} launches for the step after; identical to token-by-token")
    finally:
        a.free()
        b.free()


@pytest.mark.parametrize("path", [0, 1], ids=["cuda-core-columns", "tcgen05"])
@pytest.mark.parametrize("shape", [(4096, 4096), (1000, 4096), (4096, 11008), (260, 128)])
def test_matmul_columns_bit_exact(oracle_lib, shape, path):
    """Kernel-level: out[N][M] = W x N columns must equal ggml_compute_forward_mul_mat_q4_0_f32 bit for bit for every
    column (BASELINE.json configs[4]: bs in {1, 4, 16}, plus a ragged 33), incl. M that is not a multiple of the 128-row tile."""
    from llama_swift_b200 import ggml_format as gf
    M, K = shape
    rng = np.random.default_rng(M + K)
    wq = gf.quantize_q4_0((rng.standard_normal((M, K)) / np.sqrt(K)).astype(np.float32))
    for N in (1, 4, 16, 33):
        x = (rng.standard_normal((N, K)) * 1.3).astype(np.float32)
        x[0, :32] = 0.0
        want = np.zeros((N, M), np.float32)
        oracle_lib.ora_mul_mat_q4(2, wq.ctypes.data, M, K, x.ctypes.data, N, want.ctypes.data)
        got = lsb.q4_0_matmul(wq, x, path=path)
        assert np.array_equal(bits(got), bits(want)), f"N={N}: max abs diff {np.abs(got - want).max()}"
