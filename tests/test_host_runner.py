"""The token loop above the C ABI (csrc/host_runner.cpp == -[LlamaPredictOperation main], PO.mm:768-901) on the CPU:
b200_llama_run_loop with an injected evaluator against a run assembled HERE, independently, from the reference's own
pieces -- its tokenizer, its sampler, its llama_eval (oracle/_ref) -- following PO.mm line by line.  Same prompt, seed
and parameters must give the same emitted ids (prompt echo + sampled tokens) and the same event order."""
import ctypes as C
import os

import numpy as np
import pytest

import llama_swift_b200 as lsb
from test_host_text import N_VOCAB, _pieces, ref_model  # noqa: F401  (fixture + vocabulary)

N_CTX = 8          # what the ref_model fixture loads with


def _reference_run(L, h, prompt: bytes, p: lsb.RunParams, n_ctx: int):
    """PO.mm:768-901 restated with the reference's own functions (test-side, Python)."""
    events = [lsb.EVENT_STARTED_GENERATING_OUTPUT]
    s = C.c_void_p(L.ref_sampler_new(p.seed))
    assert prompt, "the empty-prompt path draws from the generator first; covered separately"
    buf = np.empty(len(prompt) + 2, np.int32)
    embd_inp = list(buf[:L.ref_tokenize(h, prompt, 1, buf.ctypes.data, len(buf))])
    n_predict = min(p.n_predict, n_ctx - len(embd_inp))                                  # PO.mm:812
    logits = np.empty(N_VOCAB, np.float32)
    err = C.create_string_buffer(256)
    probe = np.array([0, 1, 2, 3], np.int32)
    assert L.ref_llama_eval(h, p.n_threads, 0, probe.ctypes.data, 4, logits.ctypes.data, err, 256) == 0   # PO.mm:822
    last = [0] * p.repeat_last_n
    embd, n_past, remaining, consumed, out = [], 0, n_predict, 0, []
    while remaining > 0:
        if embd:
            t = np.array(embd, np.int32)
            assert L.ref_llama_eval(h, p.n_threads, n_past, t.ctypes.data, len(t), logits.ctypes.data, err, 256) == 0
        n_past += len(embd)
        embd = []
        if len(embd_inp) <= consumed:
            la = np.array(last, np.int32)
            f32 = np.float32
            tok = L.ref_sample_top_p_top_k(h, s, logits.ctypes.data, la.ctypes.data, len(la), float(f32(p.repeat_penalty)),
                                           int(f32(p.top_k)), float(f32(p.top_p)), float(f32(p.temp)))
            last = last[1:] + [tok]
            embd.append(tok)
            remaining -= 1
        else:
            while len(embd_inp) > consumed:
                embd.append(int(embd_inp[consumed]))
                last = last[1:] + [int(embd_inp[consumed])]
                consumed += 1
                if len(embd) > p.n_batch:
                    break
        out += embd
        events += [lsb.EVENT_OUTPUT_TOKEN] * len(embd)
    events.append(lsb.EVENT_COMPLETED)
    L.ref_sampler_free(s)
    return out, events


@pytest.mark.parametrize("prompt,overrides", [
    (b"the abc", dict(n_predict=3)),                                  # 4 prompt tokens + 3 sampled in an 8-token context
    (b"ab", dict(n_predict=100, seed=7)),                              # n_predict clamped by n_ctx - prompt (PO.mm:812)
    (b"a", dict(n_predict=5, top_k=1)),                                # greedy
    (b" w1 w2 w3", dict(n_predict=2, n_batch=1, temp=1.5, top_p=0.5)),  # prompt forwarded 2 tokens at a time
])
def test_run_loop_matches_reference_pieces(ref_lib, ref_model, prompt, overrides):
    L, h = ref_model
    p = lsb.default_run_params(n_ctx=N_CTX, **overrides)
    want_ids, want_events = _reference_run(L, h, prompt, p, N_CTX)

    # the product loop, its evaluator injected: a SECOND instance of the reference model (so the two runs share no KV state)
    err = C.create_string_buffer(256)
    path = os.path.join(os.environ.get("B200_TEST_CACHE", "/tmp/b200_llama_test_models"), "ggml-model-hosttext-v%d.bin" % N_VOCAB)
    h2 = C.c_void_p(L.ref_llama_load(os.fsencode(path), N_CTX, err, 256))
    assert h2

    def ev(_ctx, n_threads, n_past, toks, n, logits, e, elen):
        return L.ref_llama_eval(h2, n_threads, n_past, C.cast(toks, C.c_void_p), n, C.cast(logits, C.c_void_p), C.cast(e, C.c_char_p), elen)

    got_ids, got_events, pieces_seen = [], [], []

    def on_event(_user, kind, text, n, code):
        got_events.append(kind)
        if kind == lsb.EVENT_OUTPUT_TOKEN:
            got_ids.append(code)
            pieces_seen.append(C.string_at(text, n))

    pieces = _pieces()
    tok = lsb.Tokenizer(pieces=pieces)
    arr = (C.c_char_p * len(pieces))(*pieces)
    lens = (C.c_int * len(pieces))(*[len(x) for x in pieces])
    ev_c, on_c = lsb.EVAL_FN(ev), lsb.EVENT_FN(on_event)
    rc = lsb.lib().b200_llama_run_loop(ev_c, None, N_VOCAB, N_CTX, tok._h, arr, lens, prompt, len(prompt), b"", 0, C.byref(p), on_c, None)
    L.ref_llama_free(h2)
    assert rc == 0
    assert got_ids == [int(x) for x in want_ids], (got_ids, want_ids)
    assert got_events == want_events
    assert pieces_seen == [pieces[i] for i in got_ids]


def test_run_loop_reports_eval_failure():
    """A failing evaluator ends the run with a FAILED event carrying LlamaErrorCodePredictionFailed (PO.mm:823, 841)."""
    p = lsb.default_run_params(n_ctx=16, n_predict=4)
    pieces = [b"", b"", b"", b"a", b"b"]
    tok = lsb.Tokenizer(pieces=pieces)
    arr = (C.c_char_p * len(pieces))(*pieces)
    lens = (C.c_int * len(pieces))(*[len(x) for x in pieces])
    calls = []

    def ev(_ctx, n_threads, n_past, toks, n, logits, e, elen):
        calls.append((n_past, n))
        if len(calls) == 1:
            for i in range(5):
                logits[i] = float(i)
            return 0
        msg = b"boom"
        C.memmove(e, msg + b"\\0", len(msg) + 1)
        return lsb.ERR_PREDICT

    events = []

    def on_event(_user, kind, text, n, code):
        events.append((kind, C.string_at(text, n) if text else b"", code))

    ev_c, on_c = lsb.EVAL_FN(ev), lsb.EVENT_FN(on_event)
    rc = lsb.lib().b200_llama_run_loop(ev_c, None, 5, 16, tok._h, arr, lens, b"ab", 2, b"", 0, C.byref(p), on_c, None)
    assert rc == lsb.ERR_PREDICT
    assert calls[0] == (0, 4) and calls[1] == (0, 3)                   # probe, then BOS + 2 prompt tokens
    assert events[-1][0] == lsb.EVENT_FAILED and events[-1][2] == lsb.ERR_PREDICT and events[-1][1].startswith(b"boom")
