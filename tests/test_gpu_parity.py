"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same seeded inputs.

Bar (BASELINE.json north_star): logits within 1e-3 relative and identical arg-max vs the reference CPU path.  The
kernels mirror the reference's AVX2 arithmetic operation for operation, so these tests additionally demand
bit-identical results wherever the only freedom is summation order of exact quantities, and report how close the rest
is (the LayerNorm double sums are tree-ordered on the GPU)."""
import os

import numpy as np
import pytest

import llama_swift_b200 as lsb
from conftest import CpuModel, bits, model_file, rel_l2

pytestmark = pytest.mark.gpu

RNG = np.random.default_rng(99)


def _qweights(M, K, scale=None):
    from llama_swift_b200 import ggml_format as gf
    w = (RNG.standard_normal((M, K)) * (scale or 1.0 / np.sqrt(K))).astype(np.float32)
    return gf.quantize_q4_0(w)


@pytest.mark.parametrize("lp", [1, 2, 4])
@pytest.mark.parametrize("shape", [(4096, 4096), (4096, 11008), (1000, 4096), (4, 64), (6, 128), (12288, 4096)])
def test_matvec_bit_exact(oracle_lib, shape, lp):
    """mul_mat_q4_0_f32 (activation quantize + vec_dot_q4_0) must be bit-identical: integer block dots are exact and
    the fp32 lane accumulation order is the reference's."""
    M, K = shape
    wq = _qweights(M, K)
    x = (RNG.standard_normal(K) * 1.7).astype(np.float32)
    x[:32] = 0.0
    want = np.zeros(M, np.float32)
    oracle_lib.ora_mul_mat_q4(2, wq.ctypes.data, M, K, x.ctypes.data, 1, want.ctypes.data)
    got = lsb.q4_0_matvec(wq, x, lane_pairs=lp)
    assert np.array_equal(bits(got), bits(want)), f"max abs diff {np.abs(got - want).max()}"


def test_matvec_extremes(oracle_lib):
    """nibble extremes (q = 0 -> -8, q = 15 -> +7), zero scales, huge and tiny activations."""
    M, K = 64, 256
    blocks = np.zeros((M, K // 32, 20), np.uint8)
    blocks[:, :, 4:] = RNG.choice(np.array([0x00, 0xFF, 0x0F, 0xF0, 0x88], np.uint8), size=(M, K // 32, 16))
    d = (RNG.standard_normal((M, K // 32)) * 0.1).astype(np.float32)
    d[::7] = 0.0
    blocks[:, :, :4] = d.view(np.uint8).reshape(M, K // 32, 4)
    for x in ((RNG.standard_normal(K) * 1e20).astype(np.float32), (RNG.standard_normal(K) * 1e-20).astype(np.float32),
              np.zeros(K, np.float32)):
        want = np.zeros(M, np.float32)
        oracle_lib.ora_mul_mat_q4(2, blocks.ctypes.data, M, K, x.ctypes.data, 1, want.ctypes.data)
        got = lsb.q4_0_matvec(blocks, x)
        assert np.array_equal(bits(got), bits(want))


def _report(tag, got, want):
    r = rel_l2(got, want)
    same = np.array_equal(bits(got), bits(want))
    print(f"[parity] {tag}: bit-identical={same} rel_l2={r:.3e} argmax {int(got.argmax())} vs {int(want.argmax())}")
    return same, r


@pytest.mark.parametrize("mega", [1, 0], ids=["whole-token-kernel", "per-matrix-kernels"])
@pytest.mark.parametrize("n_threads", [1, 8])
def test_llama_eval_vs_oracle(oracle_lib, small_model, n_threads, mega):
    """llama_eval through the C ABI: prompt batches of 4 and 9 tokens, then single-token steps (PO.mm:822-889).
    Both schedulers are checked: the persistent whole-token kernel (default) and the per-matrix kernel sequence."""
    ora = CpuModel(oracle_lib, "ora", small_model, 64)
    gpu = lsb.llama_model_load(small_model, n_ctx=64)
    gpu.set_option("mega", mega)
    try:
        rng = np.random.default_rng(7)
        n_past, n_exact = 0, 0
        steps = (4, 9, 1, 1, 1, 1, 1, 1, 1, 1)
        for n in steps:
            toks = rng.integers(3, 512, size=n).astype(np.int32)
            want = ora.eval(n_threads, n_past, toks)
            got = lsb.llama_eval(gpu, n_threads, n_past, toks)
            same, r = _report(f"nth={n_threads} n_past={n_past} N={n}", got, want)
            assert r <= 1e-3, "logits must be within 1e-3 relative of the reference CPU path"
            assert got.argmax() == want.argmax()
            n_exact += same
            n_past += n
        for il in range(2):
            for which in (0, 1):
                g, w = gpu.kv_export(il, which, n_past), ora.kv(il, which, n_past)
                assert rel_l2(g, w) <= 1e-4
        print(f"[parity] {n_exact}/{len(steps)} steps bit-identical")
        assert n_exact >= len(steps) - 1, "expected bit-identical logits (LayerNorm tree-sum flips are ~1e-9 rare)"
        # probe pattern (PO.mm:822): n_past = 0 again over a dirty cache
        toks = np.array([0, 1, 2, 3], dtype=np.int32)
        got, want = lsb.llama_eval(gpu, n_threads, 0, toks), ora.eval(n_threads, 0, toks)
        assert rel_l2(got, want) <= 1e-3 and got.argmax() == want.argmax()
    finally:
        ora.free()
        gpu.free()


def test_decode_device_teacher_forced(oracle_lib, small_model):
    """The device-resident loop (CUDA-graph replay, scalars in HBM) against per-step oracle evals."""
    n_steps = 24
    ora = CpuModel(oracle_lib, "ora", small_model, 64)
    gpu = lsb.llama_model_load(small_model, n_ctx=64)
    try:
        rng = np.random.default_rng(3)
        stream = rng.integers(3, 512, size=n_steps + 1).astype(np.int32)
        toks, logits, ms = gpu.decode_device(0, int(stream[0]), n_steps, n_threads=8, forced_tokens=stream[1:], want_logits=True)
        worst = 0.0
        for i in range(n_steps):
            want = ora.eval(8, i, stream[i:i + 1])
            worst = max(worst, rel_l2(logits[i], want))
            assert int(toks[i]) == int(want.argmax())
        print(f"[parity] teacher-forced {n_steps} steps: worst rel_l2 {worst:.3e}, {ms:.3f} ms total")
        assert worst <= 1e-3
        # graph replay and plain launches agree bit for bit; so do the two schedulers
        gpu.set_option("graph", 0)
        toks2, logits2, _ = gpu.decode_device(0, int(stream[0]), n_steps, n_threads=8, forced_tokens=stream[1:], want_logits=True)
        assert np.array_equal(bits(logits), bits(logits2)) and np.array_equal(toks, toks2)
        gpu.set_option("mega", 0)
        toks3, logits3, _ = gpu.decode_device(0, int(stream[0]), n_steps, n_threads=8, forced_tokens=stream[1:], want_logits=True)
        assert np.array_equal(bits(logits), bits(logits3)) and np.array_equal(toks, toks3)
    finally:
        ora.free()
        gpu.free()


def test_greedy_tokens_match(oracle_lib, small_model):
    ora = CpuModel(oracle_lib, "ora", small_model, 64)
    gpu = lsb.llama_model_load(small_model, n_ctx=64)
    try:
        toks, _, _ = gpu.decode_device(0, 1, 16, n_threads=8)
        cur, want = 1, []
        for i in range(16):
            cur = int(ora.eval(8, i, np.array([cur], np.int32)).argmax())
            want.append(cur)
        assert list(map(int, toks)) == want
    finally:
        ora.free()
        gpu.free()


def test_error_behaviour(tmp_path, small_model):
    with pytest.raises(lsb.LlamaError) as ei:
        lsb.llama_model_load(str(tmp_path / "nope.bin"))
    assert ei.value.code == lsb.ERR_LOAD and "failed to open" in ei.value.message          # PO.mm:100-104
    bad = tmp_path / "bad.bin"
    bad.write_bytes(b"\x00" * 64)
    with pytest.raises(lsb.LlamaError) as ei:
        lsb.llama_model_load(str(bad))
    assert ei.value.code == lsb.ERR_LOAD and "bad magic" in ei.value.message                # PO.mm:110-114
    trunc = tmp_path / "trunc.bin"
    trunc.write_bytes(open(small_model, "rb").read(40_000_000))
    with pytest.raises(lsb.LlamaError) as ei:
        lsb.llama_model_load(str(trunc))
    assert ei.value.code == lsb.ERR_LOAD
    gpu = lsb.llama_model_load(small_model, n_ctx=16)
    try:
        with pytest.raises(lsb.LlamaError) as ei:
            lsb.llama_eval(gpu, 8, 10, np.arange(3, 12, dtype=np.int32))                     # 10 + 9 > n_ctx
        assert ei.value.code == lsb.ERR_PREDICT
        with pytest.raises(lsb.LlamaError):
            lsb.llama_eval(gpu, 8, 0, np.array([100000], np.int32))
        assert gpu.id_to_token(5) == b" t5" and gpu.id_to_token(0) == b""
    finally:
        gpu.free()


@pytest.mark.parametrize("nth", [1, 8])
def test_llama_eval_vs_golden_reference_vectors(nth):
    """The CUDA path against the committed outputs of the UNMODIFIED reference (tests/golden/reference_vectors.npz)."""
    from conftest import ROOT
    G = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))
    gpu = lsb.llama_model_load(model_file(n_layer=1, n_vocab=256, seed=11), n_ctx=16)
    try:
        toks, n_past = G["eval_tokens"], 0
        for i, n in enumerate((4, 1, 1)):
            got = lsb.llama_eval(gpu, nth, n_past, toks[n_past:n_past + n])
            want = G[f"eval_logits_nth{nth}_{i}"]
            assert rel_l2(got, want) <= 1e-3 and got.argmax() == want.argmax()
            assert np.array_equal(bits(got), bits(want)), "expected bit-identical logits vs the reference CPU run"
            n_past += n
        assert np.array_equal(bits(gpu.kv_export(0, 0, n_past)[:, :256]), bits(G[f"eval_k0_nth{nth}"]))
        assert np.array_equal(bits(gpu.kv_export(0, 1, n_past)[:, :256]), bits(G[f"eval_v0_nth{nth}"]))
    finally:
        gpu.free()


@pytest.mark.parametrize("shape", [(4096, 4096), (1000, 4096), (4096, 11008), (8, 64)])
def test_matvec_q4_1_bit_exact(oracle_lib, shape):
    """mul_mat_q4_1_f32: quantize_row_q4_1 + the scalar, strictly sequential vec_dot_q4_1 chain (ggml.c:1584-1626)."""
    from llama_swift_b200 import ggml_format as gf
    M, K = shape
    w = (RNG.standard_normal((M, K)) / np.sqrt(K)).astype(np.float32)
    wq = gf.quantize_q4_1(w)
    x = (RNG.standard_normal(K) * 1.7).astype(np.float32)
    x[:32] = 0.25                                   # constant block: d = 0, id = 0
    want = np.zeros(M, np.float32)
    oracle_lib.ora_mul_mat_q4(3, wq.ctypes.data, M, K, x.ctypes.data, 1, want.ctypes.data)
    got = lsb.q4_1_matvec(wq, x)
    assert np.array_equal(bits(got), bits(want)), f"max abs diff {np.abs(got - want).max()}"


def test_llama_eval_q4_1_vs_oracle(oracle_lib):
    """A Q4_1 model file (type 3, quantized by the restated utils.cpp quantizer) through llama_eval: bit-identical."""
    path = model_file(n_layer=1, n_vocab=256, seed=21, ftype=3)
    ora = CpuModel(oracle_lib, "ora", path, 32)
    gpu = lsb.llama_model_load(path, n_ctx=32)
    try:
        assert gpu.ftype == 3
        rng = np.random.default_rng(5)
        n_past = 0
        for n in (4, 1, 1, 1):
            toks = rng.integers(3, 256, size=n).astype(np.int32)
            want = ora.eval(8, n_past, toks)
            got = lsb.llama_eval(gpu, 8, n_past, toks)
            same, r = _report(f"q4_1 n_past={n_past} N={n}", got, want)
            assert r <= 1e-3 and got.argmax() == want.argmax()
            assert same
            n_past += n
    finally:
        ora.free()
        gpu.free()


def test_13b_shapes_vs_oracle(oracle_lib):
    """LLaMA-13B geometry (n_embd 5120, 40 heads, n_ff 13824; BASELINE.json configs[3]) from a TWO-PART model file
    (LLAMA_N_PARTS, PO.mm:33-38; column/row merge PO.mm:358-388): non-power-of-two LayerNorm length, 4 quantization rounds
    for the w2 input, 188-row CTAs.  One layer, both schedulers."""
    path = model_file(n_layer=1, n_vocab=512, seed=31, n_embd=5120)
    assert os.path.exists(path + ".1"), "expected a two-part model file"
    ora = CpuModel(oracle_lib, "ora", path, 32)
    try:
        for mega in (1, 0):
            gpu = lsb.llama_model_load(path, n_ctx=32)
            gpu.set_option("mega", mega)
            try:
                assert gpu.n_embd == 5120 and gpu.n_head == 40
                rng = np.random.default_rng(9)
                n_past, exact = 0, 0
                steps = (4, 9, 1, 1, 1)
                for n in steps:
                    toks = rng.integers(3, 512, size=n).astype(np.int32)
                    want = ora.eval(8, n_past, toks)
                    got = lsb.llama_eval(gpu, 8, n_past, toks)
                    same, r = _report(f"13B-shape mega={mega} n_past={n_past} N={n}", got, want)
                    assert r <= 1e-3 and got.argmax() == want.argmax()
                    exact += same
                    n_past += n
                assert exact >= len(steps) - 1
            finally:
                gpu.free()
    finally:
        ora.free()


def test_resident_model_cache(oracle_lib, small_model):
    """the reference re-reads the model file on every run() (PO.mm:790); an acquired model that
    was released is handed back without loading, and a second run over its dirty KV cache gives the same logits."""
    import time
    lsb.llama_model_cache_clear()
    toks = np.array([1, 17, 33, 250, 9], np.int32)
    t0 = time.perf_counter()
    a = lsb.llama_model_acquire(small_model, n_ctx=64)
    t_load = time.perf_counter() - t0
    first = lsb.llama_eval(a, 8, 0, toks)
    step = lsb.llama_eval(a, 8, 5, np.array([int(first.argmax())], np.int32))
    h1 = a._h.value
    busy = lsb.llama_model_acquire(small_model, n_ctx=64)       # `a` is in use: a concurrent operation gets its own model
    assert busy._h.value != h1
    h2 = busy._h.value
    busy.release()
    a.release()
    t0 = time.perf_counter()
    b = lsb.llama_model_acquire(small_model, n_ctx=64)
    t_again = time.perf_counter() - t0
    try:
        assert b._h.value in (h1, h2)                              # a resident model (one idle copy is kept), not a fresh load
        again = lsb.llama_eval(b, 8, 0, toks)
        assert np.array_equal(bits(again), bits(first))
        ora = CpuModel(oracle_lib, "ora", small_model, 64)            # ... and both are the oracle's logits
        try:
            assert np.array_equal(bits(again), bits(ora.eval(8, 0, toks)))
        finally:
            ora.free()
        assert np.array_equal(bits(lsb.llama_eval(b, 8, 5, np.array([int(first.argmax())], np.int32))), bits(step))
        print(f"[cache] first acquire {t_load * 1e3:.0f} ms, re-acquire {t_again * 1e3:.2f} ms")
        assert t_again < 0.05 * t_load + 0.005
        c = lsb.llama_model_acquire(small_model, n_ctx=32)         # another n_ctx is another model
        assert c.n_ctx == 32
        c.release()
    finally:
        b.release()
        lsb.llama_model_cache_clear()


@pytest.mark.parametrize("n_threads", [8, 3])
def test_long_context_teacher_forced(oracle_lib, small_model, n_threads):
    """Positions up to 280: several K.Q rounds per warp, V.P chains long enough for every batch size of the chain loop
    (n_threads 3 -> 94 positions per chain) and the L2 prefetch of earlier KV rows.  Bit-identical at every step."""
    n_steps = 280
    ora = CpuModel(oracle_lib, "ora", small_model, 288)
    gpu = lsb.llama_model_load(small_model, n_ctx=288)
    try:
        rng = np.random.default_rng(17)
        stream = rng.integers(3, 512, size=n_steps + 1).astype(np.int32)
        toks, logits, _ = gpu.decode_device(0, int(stream[0]), n_steps, n_threads=n_threads, forced_tokens=stream[1:], want_logits=True)
        exact, worst = 0, 0.0
        for i in range(n_steps):
            want = ora.eval(n_threads, i, stream[i:i + 1])
            worst = max(worst, rel_l2(logits[i], want))
            assert int(toks[i]) == int(want.argmax()), i
            exact += int(np.array_equal(bits(logits[i]), bits(want)))
        print(f"[parity] long context nth={n_threads}: {exact}/{n_steps} steps bit-identical, worst rel_l2 {worst:.3e}")
        assert worst <= 1e-3 and exact >= n_steps - 2
        for il in range(2):
            assert np.array_equal(bits(gpu.kv_export(il, 1, n_steps)), bits(ora.kv(il, 1, n_steps)))
    finally:
        ora.free()
        gpu.free()


def test_bounded_wait_abort_and_recover(oracle_lib, small_model):
    """a device-side wait that times out must not poison the CUDA context.  With a 300-tick wait budget
    the token kernel gives up (abort word, csrc/ptx.cuh), the call returns -1001 (the reference's predict error code,
    LlamaError.h:18), the exchange state is reset, and the NEXT call is bit-identical to the oracle again."""
    ora = CpuModel(oracle_lib, "ora", small_model, 64)
    gpu = lsb.llama_model_load(small_model, n_ctx=64)
    try:
        toks = np.array([1, 17, 33, 250, 9], np.int32)
        gpu.set_option("batch", 0)
        want = ora.eval(8, 0, toks)
        assert np.array_equal(bits(lsb.llama_eval(gpu, 8, 0, toks)), bits(want))
        gpu.set_option("spin_limit_cycles", 300)
        with pytest.raises(lsb.LlamaError) as ei:
            for _ in range(4):      # some launch will have a wait longer than 300 ticks
                lsb.llama_eval(gpu, 8, 0, toks)
        assert ei.value.code == -1001
        gpu.set_option("spin_limit_cycles", 0)
        again = lsb.llama_eval(gpu, 8, 0, toks)
        assert np.array_equal(bits(again), bits(want))
        step = lsb.llama_eval(gpu, 8, 5, np.array([int(want.argmax())], np.int32))
        assert np.array_equal(bits(step), bits(ora.eval(8, 5, np.array([int(want.argmax())], np.int32))))
    finally:
        ora.free()
        gpu.free()
