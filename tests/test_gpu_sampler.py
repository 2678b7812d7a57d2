"""The sampler's candidate stage on the GPU (sample_topk.cuh) against llama_sample_top_p_top_k itself.

The reference function (utils.cpp:345-428, compiled into oracle/_ref) is run with a fresh std::mt19937 of the same seed on the
same logits; the GPU path (candidates on the device, draw on the host) must return the same id -- and must say "ambiguous"
exactly when two of the best top_k + 1 values compare equal, which is when only the reference's own partial_sort knows the order."""
import ctypes as C

import numpy as np
import pytest

import llama_swift_b200 as lsb
from test_host_text import ref_model  # noqa: F401

pytestmark = pytest.mark.gpu


def _host_candidates(logits, last, penalty, top_k, temp):
    """utils.cpp:357-386 in numpy doubles (values only; the tests below use it where all values are distinct)."""
    scale = 1.0 / np.float64(temp)
    v = logits.astype(np.float64) * scale
    ids = np.unique(last[(last >= 0) & (last < len(logits))])
    neg = logits[ids] < 0
    v[ids] = np.where(neg, logits[ids].astype(np.float64) * scale * np.float64(penalty), logits[ids].astype(np.float64) * scale / np.float64(penalty))
    order = np.argsort(-v, kind="stable")[:top_k]
    return v[order], order.astype(np.int32)


@pytest.mark.parametrize("n_vocab", [32000, 512, 33, 40000])
@pytest.mark.parametrize("top_k,temp,penalty", [(40, 0.8, 1.3), (1, 1.0, 1.0), (64, 0.1, 2.5)])
def test_candidates_match_reference_formula(n_vocab, top_k, temp, penalty):
    rng = np.random.default_rng(n_vocab + top_k)
    for trial in range(4):
        logits = (rng.standard_normal(n_vocab) * 4.0).astype(np.float32)
        last = rng.integers(0, n_vocab, size=64).astype(np.int32)
        last[:8] = np.argsort(-logits)[:8]              # the best tokens are in the window (the usual case while generating)
        last[8] = 0
        got = lsb.sample_topk(logits, last, penalty, top_k, temp)
        want_v, want_i = _host_candidates(logits, last, penalty, min(top_k, n_vocab), temp)
        assert got is not None
        assert np.array_equal(got[1], want_i)
        assert np.array_equal(got[0].view(np.uint64), want_v.view(np.uint64))


def test_ambiguous_orders_are_reported():
    rng = np.random.default_rng(7)
    logits = (rng.standard_normal(32000) * 4.0).astype(np.float32)
    last = np.zeros(64, np.int32)
    top = np.argsort(-logits)
    # two equal values inside the best 40
    a = logits.copy(); a[top[5]] = a[top[6]]
    assert lsb.sample_topk(a, last) is None
    # equal values across the cut (40th and 41st)
    b = logits.copy(); b[top[40]] = b[top[39]]
    assert lsb.sample_topk(b, last) is None
    # equal values below the cut do not matter
    c = logits.copy(); c[top[100]] = c[top[101]]
    assert lsb.sample_topk(c, last) is not None
    # all logits equal / NaN
    assert lsb.sample_topk(np.zeros(32000, np.float32), last) is None
    d = logits.copy(); d[top[0]] = np.nan
    assert lsb.sample_topk(d, last) is None


def _draws(n_vocab, n_steps, reference_draw, seed=1234):
    """n_steps sampling steps on a synthetic logits stream; returns how many were served with the candidate stage on the GPU."""
    rng = np.random.default_rng(11)
    ours = lsb.Sampler(seed)
    last = np.zeros(64, np.int32)                                      # PO.mm:828-829
    pen, top_p, temp = float(np.float32(1.3)), float(np.float32(0.95)), float(np.float32(0.8))   # PO.mm:852-855: const float locals
    n_gpu = 0
    for step in range(n_steps):
        logits = (rng.standard_normal(n_vocab) * (1.0 + step % 5)).astype(np.float32)
        if step % 50 == 49:
            logits[:] = np.round(logits)                               # heavy ties: the fallback must kick in and still match
        want = reference_draw(logits, last, pen, 40, top_p, temp)
        cand = lsb.sample_topk(logits, last, pen, 40, temp)
        if cand is not None:
            got = ours.sample_from_candidates(cand[0], cand[1], top_p)
            n_gpu += 1
        else:
            got = ours.sample(logits, last, pen, 40, top_p, temp)
        assert got == want, f"step {step}"
        last = np.roll(last, -1)                                       # PO.mm:867-868
        last[-1] = want
    return n_gpu


def test_ids_identical_to_reference_sampler(ref_model):
    """Same seed, same logits stream: the reference's own llama_sample_top_p_top_k (compiled from /root/reference into oracle/_ref,
    vocabulary of the fixture model) vs GPU candidates + host draw."""
    from test_host_text import N_VOCAB
    L, h = ref_model
    ref_s = C.c_void_p(L.ref_sampler_new(1234))
    try:
        n_gpu = _draws(N_VOCAB, 200, lambda lg, last, pen, k, tp, t: L.ref_sample_top_p_top_k(h, ref_s, lg.ctypes.data, last.ctypes.data, len(last), pen, k, tp, t))
    finally:
        L.ref_sampler_free(ref_s)
    assert n_gpu >= 190
    print(f"[sampler] n_vocab {N_VOCAB}: 200 draws identical to the reference; {n_gpu} with the candidate stage on the GPU")


def test_ids_identical_at_llama_vocab():
    """n_vocab 32000: against the host sampler (itself pinned to the reference in tests/test_host_text.py)."""
    host = lsb.Sampler(1234)
    n_gpu = _draws(32000, 120, lambda lg, last, pen, k, tp, t: host.sample(lg, last, pen, k, tp, t))
    assert n_gpu >= 114
    print(f"[sampler] n_vocab 32000: 120 draws identical to the host sampler; {n_gpu} with the candidate stage on the GPU")


def test_candidate_kernel_time():
    rng = np.random.default_rng(3)
    logits = (rng.standard_normal(32000) * 4.0).astype(np.float32)
    last = rng.integers(0, 32000, size=64).astype(np.int32)
    res, ms = lsb.sample_topk(logits, last, timed=True)
    assert res is not None
    print(f"[sampler] candidate kernel, n_vocab 32000, top_k 40: {ms * 1e3:.1f} us")
    assert ms < 0.2
