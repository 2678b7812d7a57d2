"""Host-side callees of the reference's token loop: the tokenizer and the sampler of
include/b200_llama.h against the UNMODIFIED reference functions (utils.cpp, compiled into oracle/_ref) -- identical ids.
CPU only: no GPU is involved in either."""
import ctypes as C
import os
import time

import numpy as np
import pytest

import llama_swift_b200 as lsb
from llama_swift_b200 import ggml_format as gf
from conftest import CACHE

N_VOCAB = 320


def _pieces():
    """A nasty little vocabulary: empty control pieces, shared prefixes, a duplicate, multi-byte UTF-8, high bytes."""
    base = [b"", b"", b"", b"a", b"ab", b"abc", b"abcd", b"b", b"bc", b"c", b"ab", b" ", b" the", b" th", b"the", b"t", b"h", b"e",
            "é".encode(), "éa".encode(), "日本".encode(), "日".encode(), b"\xff", b"\xff\xfe", b"xyzxyzxyzxyzxyzxyz", b"xyz", b"x", b"y", b"z",
            b"\n", b"\n\n", b"0", b"00", b"000", b"1"]
    out = list(base)
    i = 0
    while len(out) < N_VOCAB:
        out.append((" w%d" % i).encode())
        i += 1
    return out


@pytest.fixture(scope="module")
def ref_model(ref_lib):
    os.makedirs(CACHE, exist_ok=True)
    path = os.path.join(CACHE, "ggml-model-hosttext-v%d.bin" % N_VOCAB)
    if not os.path.exists(path):
        gf.write_synthetic_model(path + ".tmp", gf.HParams(n_vocab=N_VOCAB, n_layer=1), seed=3, mode="direct", vocab=_pieces())
        os.replace(path + ".tmp", path)
    L = ref_lib
    vp, ci, cp = C.c_void_p, C.c_int, C.c_char_p
    L.ref_tokenize.argtypes, L.ref_tokenize.restype = [vp, cp, ci, vp, ci], ci
    L.ref_sampler_new.argtypes, L.ref_sampler_new.restype = [ci], vp
    L.ref_sampler_free.argtypes = [vp]
    L.ref_sample_top_p_top_k.argtypes = [vp, vp, vp, vp, ci, C.c_double, ci, C.c_double, C.c_double]
    L.ref_sample_top_p_top_k.restype = ci
    err = C.create_string_buffer(256)
    h = L.ref_llama_load(os.fsencode(path), 8, err, 256)
    assert h, err.value
    yield L, C.c_void_p(h)
    L.ref_llama_free(C.c_void_p(h))


def _ref_tokens(L, h, text: bytes, bos: bool):
    out = np.empty(len(text) + 2, np.int32)
    n = L.ref_tokenize(h, text, int(bos), out.ctypes.data, len(out))
    return out[:n]


def test_tokenizer_matches_reference(ref_model):
    L, h = ref_model
    pieces = _pieces()
    tok = lsb.Tokenizer(pieces=pieces)
    rng = np.random.default_rng(0)
    texts = [b"", b"a", b"abcd", b"abcabx", b"the theth e", "日本日é éa".encode(), b"\xff\xfe\xff", b"q", b"abq rest is dropped",
             b"xyzxyzxyzxyzxyzxyzxyz", b"000000010", b" w1 w12 w123 w2"]
    nonempty = [p for p in pieces if p]
    for _ in range(300):        # random concatenations of pieces, sometimes with a byte no piece starts with
        parts = [nonempty[i] for i in rng.integers(0, len(nonempty), size=rng.integers(1, 40))]
        if rng.random() < 0.3:
            parts.insert(int(rng.integers(0, len(parts) + 1)), b"Q")
        texts.append(b"".join(parts))
    for text in texts:
        for bos in (True, False):
            want = _ref_tokens(L, h, text, bos)
            got = tok(text, bos=bos)
            assert np.array_equal(got, want), (text, bos, got, want)


def test_tokenizer_small_buffer_reports_needed_length():
    tok = lsb.Tokenizer(pieces=[b"", b"", b"", b"a", b"b"])
    out = np.empty(2, np.int32)
    n = lsb.lib().b200_llama_tokenize(tok._h, b"abab", 4, 1, out.ctypes.data, 2)
    assert n == 5 and list(out) == [1, 3]


@pytest.mark.parametrize("params", [dict(), dict(top_k=1), dict(top_p=1.0), dict(temp=0.1), dict(temp=2.5, top_k=100),
                                    dict(repeat_penalty=1.0), dict(top_k=N_VOCAB, top_p=0.5)])
def test_sampler_matches_reference(ref_model, params):
    L, h = ref_model
    p = dict(repeat_penalty=1.3, top_k=40, top_p=0.95, temp=0.8)       # gpt_params defaults, utils.h:15-37
    p.update(params)
    rng = np.random.default_rng(42)
    for seed in (-1, 0, 12345):
        ref_s = C.c_void_p(L.ref_sampler_new(seed))
        mine = lsb.Sampler(seed)
        last = np.zeros(64, np.int32)                                    # PO.mm:828-829
        for step in range(60):
            logits = (rng.standard_normal(N_VOCAB) * 4.0).astype(np.float32)
            if step % 7 == 3:
                logits[rng.integers(0, N_VOCAB, size=12)] = logits.max()   # exact ties at the top
            if step % 11 == 5:
                logits[:] = np.float32(1.5)                                  # everything tied
            want = L.ref_sample_top_p_top_k(h, ref_s, logits.ctypes.data, last.ctypes.data, len(last),
                                            p["repeat_penalty"], p["top_k"], p["top_p"], p["temp"])
            got = mine.sample(logits, last, **p)
            assert got == want, (seed, step, got, want)
            last = np.roll(last, -1)                                      # PO.mm:867-868
            last[-1] = got
        L.ref_sampler_free(ref_s)


def test_sampler_speed_at_llama_vocab(ref_lib):
    """Not an assertion on speed, only a report: the reference's linear std::find per logit vs the bitmap (n_vocab 32000)."""
    n = 32000
    rng = np.random.default_rng(1)
    logits = (rng.standard_normal(n) * 4.0).astype(np.float32)
    last = rng.integers(0, n, size=64).astype(np.int32)
    s = lsb.Sampler(-1)
    t0 = time.perf_counter()
    for _ in range(20):
        s.sample(logits, last)
    dt = (time.perf_counter() - t0) / 20
    print(f"[host] b200_llama_sample_top_p_top_k at n_vocab {n}: {dt * 1e6:.0f} us per token")
    assert dt < 0.05


def _tokenize_restated(pieces, text: bytes, bos: bool):
    """llama_tokenize (utils.cpp:275-311) restated literally: scan the vocabulary in id order at every position."""
    out = [1] if bos else []
    pos = 0
    while True:
        best_len, best_id = 0, 0
        for i, p in enumerate(pieces):
            if len(p) < best_len or len(p) > len(text) - pos:
                continue
            if text[pos:pos + len(p)] == p:
                best_len, best_id = len(p), i
        if best_len == 0:
            break
        out.append(best_id)
        pos += best_len
    return out


def test_tokenizer_random_vocabularies():
    """Property test over random small vocabularies (tiny alphabets make prefixes, duplicates and dead ends frequent)."""
    from hypothesis import given, settings, strategies as st

    piece = st.binary(min_size=0, max_size=5).map(lambda b: bytes(97 + (x % 3) for x in b))     # alphabet {a, b, c}
    text = st.binary(min_size=0, max_size=40).map(lambda b: bytes(97 + (x % 4) for x in b))      # 'd' is never in a piece

    @settings(max_examples=300, deadline=None)
    @given(st.lists(piece, min_size=4, max_size=24), text, st.booleans())
    def check(pieces, t, bos):
        tok = lsb.Tokenizer(pieces=pieces)
        assert list(tok(t, bos=bos)) == _tokenize_restated(pieces, t, bos)

    check()


def test_sample_from_candidates_matches_full_sampler():
    """b200_llama_sample_from_candidates (the tail of llama_sample_top_p_top_k, utils.cpp:388-428) on the candidate list the GPU
    stage would return -- here formed on the host, utils.cpp:357-386 in numpy doubles -- draws the same ids as the full sampler."""
    n = 2000
    rng = np.random.default_rng(9)
    for p in (dict(), dict(top_k=1), dict(top_p=1.0), dict(temp=0.3, repeat_penalty=1.0), dict(top_k=64, top_p=0.5)):
        q = dict(repeat_penalty=1.3, top_k=40, top_p=0.95, temp=0.8)
        q.update(p)
        full, tail = lsb.Sampler(5), lsb.Sampler(5)
        last = np.zeros(64, np.int32)
        for step in range(40):
            logits = (rng.standard_normal(n) * 3.0).astype(np.float32)
            want = full.sample(logits, last, **q)
            v = logits.astype(np.float64) * (1.0 / q["temp"])
            ids = np.unique(last)
            v[ids] = np.where(logits[ids] < 0, logits[ids].astype(np.float64) * (1.0 / q["temp"]) * q["repeat_penalty"],
                              logits[ids].astype(np.float64) * (1.0 / q["temp"]) / q["repeat_penalty"])
            order = np.argsort(-v, kind="stable")[:q["top_k"]]
            assert len(np.unique(v[order])) == len(order)           # distinct values: the order is determined by the values alone
            got = tail.sample_from_candidates(v[order], order.astype(np.int32), q["top_p"])
            assert got == want, (p, step)
            last = np.roll(last, -1)
            last[-1] = want
