"""The whole token loop on the GPU: LlamaRunner.run (b200_llama_run == -[LlamaPredictOperation main], PO.mm:768-901) against
the same loop assembled from the reference's own tokenizer, sampler and CPU llama_eval -- identical emitted ids."""
import ctypes as C
import os

import pytest

import llama_swift_b200 as lsb
from test_host_runner import _reference_run
from test_host_text import N_VOCAB, _pieces, ref_model  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prompt,overrides", [("the abc 000", dict(n_predict=20, seed=-1)),
                                              (" w1 w2 w3 w4 w5 w6 w7 w8 w9 w10 w11", dict(n_predict=12, seed=3, temp=1.2))])
def test_runner_matches_reference_loop(ref_lib, ref_model, prompt, overrides):
    L, _ = ref_model
    path = os.path.join(os.environ.get("B200_TEST_CACHE", "/tmp/b200_llama_test_models"), "ggml-model-hosttext-v%d.bin" % N_VOCAB)
    n_ctx = 64
    err = C.create_string_buffer(256)
    h = C.c_void_p(L.ref_llama_load(os.fsencode(path), n_ctx, err, 256))
    assert h, err.value
    p = lsb.default_run_params(n_ctx=n_ctx, **overrides)
    try:
        want_ids, want_events = _reference_run(L, h, prompt.encode(), p, n_ctx)
    finally:
        L.ref_llama_free(h)
    events = []
    got = lsb.LlamaRunner(path).run(prompt, params=p, on_event=lambda kind, piece, code: events.append(kind))
    on_gpu, on_host = lsb.run_sampler_stats()
    # the same run with the whole sampler on the host (128 KB of logits per token): same ids
    os.environ["B200_HOST_SAMPLER"] = "1"
    try:
        got_host = lsb.LlamaRunner(path).run(prompt, params=p)
    finally:
        del os.environ["B200_HOST_SAMPLER"]
    assert lsb.run_sampler_stats()[0] == 0
    lsb.llama_model_cache_clear()
    assert got_host == got
    assert on_gpu + on_host == p.n_predict and on_gpu >= p.n_predict - 2, (on_gpu, on_host)
    print(f"[runner] {on_gpu} of {p.n_predict} sampling steps with the candidate stage on the GPU")
    pieces = _pieces()
    assert [i for i, _ in got] == [int(x) for x in want_ids]
    assert [s for _, s in got] == [pieces[i] for i, _ in got]
    assert events == [lsb.EVENT_STARTED_LOADING_MODEL, lsb.EVENT_FINISHED_LOADING_MODEL] + want_events


def test_runner_load_failure_event(tmp_path):
    events = []
    with pytest.raises(lsb.LlamaError) as ei:
        lsb.LlamaRunner(str(tmp_path / "missing.bin")).run("x", on_event=lambda kind, piece, code: events.append((kind, code)))
    assert ei.value.code == lsb.ERR_LOAD
    assert events[0][0] == lsb.EVENT_STARTED_LOADING_MODEL and events[-1] == (lsb.EVENT_FAILED, lsb.ERR_LOAD)
