"""Pin the oracle: the C restatement (oracle/llama_oracle.c) must be bit-identical to the UNMODIFIED reference
(ggml.c / utils.cpp / llama_eval compiled by oracle/Makefile into oracle/_ref/libllama_ref.so), op by op and for
whole llama_eval calls.  The reference has no tests or golden vectors of its own, so the
compiled reference itself is the pin."""
import ctypes as C

import numpy as np
import pytest

from conftest import CpuModel, bits

RNG = np.random.default_rng(1234)


def _x(n, scale=1.7):
    return (RNG.standard_normal(n) * scale).astype(np.float32)


@pytest.mark.parametrize("k", [64, 4096, 11008])
@pytest.mark.parametrize("qtype", [2, 3])
def test_quantize_row(oracle_lib, ref_lib, k, qtype):
    x = _x(k)
    x[:32] = 0.0                      # all-zero block: d = 0, id = 0 (ggml.c:482)
    x[32:64] = np.float32(-3.5)       # constant block
    x[64 % k] = np.float32(1e-30)
    nbytes = k // 32 * (20 if qtype == 2 else 24)
    a = np.zeros(nbytes, np.uint8)
    b = np.zeros(nbytes, np.uint8)
    (oracle_lib.ora_quantize_row_q4_0 if qtype == 2 else oracle_lib.ora_quantize_row_q4_1)(x.ctypes.data, a.ctypes.data, k)
    ref_lib.ref_quantize_row(qtype, x.ctypes.data, b.ctypes.data, k)
    assert np.array_equal(a, b)


def test_quantize_row_ties(oracle_lib, ref_lib):
    """x*id exactly on .5 boundaries: the AVX2 build rounds to nearest EVEN (ggml.c:492-495), unlike the scalar build."""
    x = np.zeros(64, np.float32)
    x[0] = 7.0                                         # amax = 7 -> id = 1
    x[1:16] = np.arange(-7, 8, dtype=np.float32)[:15] + 0.5
    x[32] = -7.0
    x[33:48] = np.arange(-7, 8, dtype=np.float32)[:15] - 0.5
    a = np.zeros(40, np.uint8)
    b = np.zeros(40, np.uint8)
    oracle_lib.ora_quantize_row_q4_0(x.ctypes.data, a.ctypes.data, 64)
    ref_lib.ref_quantize_row(2, x.ctypes.data, b.ctypes.data, 64)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("qtype", [2, 3])
@pytest.mark.parametrize("shape", [(64, 64, 1), (100, 4096, 1), (36, 11008, 3), (256, 256, 9)])
def test_mul_mat(oracle_lib, ref_lib, qtype, shape):
    M, K, N = shape
    w = (RNG.standard_normal((M, K)) / np.sqrt(K)).astype(np.float32)
    wq = np.zeros(M * K // 32 * (20 if qtype == 2 else 24), np.uint8)
    ref_lib.ref_quantize_weights(qtype, w.copy().ctypes.data, wq.ctypes.data, M * K, K)
    x = _x(N * K)
    a = np.zeros(N * M, np.float32)
    b = np.zeros(N * M, np.float32)
    oracle_lib.ora_mul_mat_q4(qtype, wq.ctypes.data, M, K, x.ctypes.data, N, a.ctypes.data)
    for nth in (1, 4):
        ref_lib.ref_mul_mat_q4(qtype, wq.ctypes.data, M, K, x.ctypes.data, N, b.ctypes.data, nth)
        assert np.array_equal(bits(a), bits(b))


def test_norm_silu_softmax_rope(oracle_lib, ref_lib):
    n = 4096
    x = _x(3 * n).reshape(3, n)
    a = np.zeros_like(x)
    b = np.zeros_like(x)
    for i in range(3):
        oracle_lib.ora_norm(x[i].ctypes.data, a[i].ctypes.data, n)
    ref_lib.ref_unary_op(0, x.ctypes.data, n, 3, 1, 0, b.ctypes.data, 1)
    assert np.array_equal(bits(a), bits(b))

    xs = _x(11008, 4.0)
    a = np.zeros_like(xs)
    b = np.zeros_like(xs)
    oracle_lib.ora_silu(xs.ctypes.data, a.ctypes.data, xs.size)
    ref_lib.ref_unary_op(1, xs.ctypes.data, xs.size, 1, 1, 0, b.ctypes.data, 1)
    assert np.array_equal(bits(a), bits(b))

    for nc in (1, 7, 201, 512):
        p = _x(nc, 3.0)
        if nc > 3:
            p[-2:] = -np.inf                        # masked tail (ggml.c:7026-7027)
        a = p.copy()
        b = np.zeros_like(p)
        oracle_lib.ora_soft_max(a.ctypes.data, nc)
        ref_lib.ref_unary_op(2, p.ctypes.data, nc, 1, 1, 0, b.ctypes.data, 1)
        assert np.array_equal(bits(a), bits(b))

    q = _x(3 * 4096).reshape(3, 32, 128)            # [rows, n_head, head_dim]
    b = np.zeros_like(q)
    ref_lib.ref_unary_op(3, q.ctypes.data, 128, 32, 3, 17, b.ctypes.data, 1)   # mode 0: row i at position 17 + i
    a = q.copy()
    for i in range(3):
        oracle_lib.ora_rope(a[i].ctypes.data, 32, 128, 17 + i)
    assert np.array_equal(bits(a), bits(b))


@pytest.mark.parametrize("n_threads", [1, 3, 8])
def test_llama_eval_bit_exact(oracle_lib, ref_lib, small_model, n_threads):
    """Whole forward passes: prompt batches (N = 4 and 9 like PO.mm:822,885) then single-token steps; logits and
    the KV cache must be bit-identical to the reference run with the same thread count."""
    ref = CpuModel(ref_lib, "ref_llama", small_model, 64)
    ora = CpuModel(oracle_lib, "ora", small_model, 64)
    try:
        rng = np.random.default_rng(7)
        n_past = 0
        for n in (4, 9, 1, 1, 1, 1, 1, 1):
            toks = rng.integers(3, 512, size=n).astype(np.int32)
            a = ora.eval(n_threads, n_past, toks)
            b = ref.eval(n_threads, n_past, toks)
            assert np.array_equal(bits(a), bits(b)), f"logits differ at n_past={n_past}"
            n_past += n
        for il in range(2):
            for which in (0, 1):
                assert np.array_equal(bits(ora.kv(il, which, n_past)), bits(ref.kv(il, which, n_past)))
        # the probe pattern of PO.mm:822: re-evaluate from n_past = 0 over an existing cache
        toks = np.array([0, 1, 2, 3], dtype=np.int32)
        assert np.array_equal(bits(ora.eval(n_threads, 0, toks)), bits(ref.eval(n_threads, 0, toks)))
    finally:
        ref.free()
        ora.free()


def test_llama_eval_q4_1_bit_exact(oracle_lib, ref_lib):
    """Type-3 (Q4_1) model files: loader + llama_eval of the restatement against the compiled reference."""
    from conftest import model_file
    path = model_file(n_layer=1, n_vocab=256, seed=21, ftype=3)
    ref = CpuModel(ref_lib, "ref_llama", path, 32)
    ora = CpuModel(oracle_lib, "ora", path, 32)
    try:
        rng = np.random.default_rng(5)
        n_past = 0
        for n in (4, 1, 1, 1):
            toks = rng.integers(3, 256, size=n).astype(np.int32)
            assert np.array_equal(bits(ora.eval(8, n_past, toks)), bits(ref.eval(8, n_past, toks)))
            n_past += n
    finally:
        ref.free()
        ora.free()
