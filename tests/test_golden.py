"""The oracle (C restatement) against the committed golden vectors: outputs of the UNMODIFIED reference recorded by
tests/golden/make_golden.py.  Runs everywhere (no /root/reference, no GPU needed)."""
import os

import numpy as np
import pytest

from conftest import CpuModel, ROOT, bits, model_file

G = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))
GOLDEN_MODEL = dict(n_layer=1, n_vocab=256, seed=11)


@pytest.mark.parametrize("qtype", [2, 3])
def test_quantize_row_golden(oracle_lib, qtype):
    x = G["q_x"]
    q = np.zeros_like(G[f"q_row_type{qtype}"])
    (oracle_lib.ora_quantize_row_q4_0 if qtype == 2 else oracle_lib.ora_quantize_row_q4_1)(x.ctypes.data, q.ctypes.data, 256)
    assert np.array_equal(q, G[f"q_row_type{qtype}"])


@pytest.mark.parametrize("qtype", [2, 3])
def test_mul_mat_golden(oracle_lib, qtype):
    from llama_swift_b200 import ggml_format as gf
    w, x = G["mm_w"], G["mm_x"]
    wq = (gf.quantize_q4_0(w) if qtype == 2 else gf.quantize_q4_1(w)).reshape(-1)
    assert np.array_equal(wq, G[f"mm_wq_type{qtype}"]), "offline quantizer restatement differs from utils.cpp"
    y = np.zeros(2 * 48, np.float32)
    oracle_lib.ora_mul_mat_q4(qtype, wq.ctypes.data, 48, 256, x.ctypes.data, 2, y.ctypes.data)
    assert np.array_equal(bits(y), bits(G[f"mm_y_type{qtype}"]))


def test_small_ops_golden(oracle_lib):
    x = G["u_x"]
    y = np.zeros(512, np.float32)
    oracle_lib.ora_norm(x.ctypes.data, y.ctypes.data, 512)
    assert np.array_equal(bits(y), bits(G["u_norm"]))
    oracle_lib.ora_silu(x.ctypes.data, y.ctypes.data, 512)
    assert np.array_equal(bits(y), bits(G["u_silu"]))
    p = x.copy()
    oracle_lib.ora_soft_max(p.ctypes.data, 512)
    assert np.array_equal(bits(p), bits(G["u_soft_max"]))
    r = x.copy()
    oracle_lib.ora_rope(r.ctypes.data, 4, 128, 5)
    assert np.array_equal(bits(r), bits(G["u_rope_p5"]))


@pytest.mark.parametrize("nth", [1, 8])
def test_llama_eval_golden(oracle_lib, nth):
    """Also pins the model writer: the file is re-created from its seed and must reproduce the recorded logits."""
    m = CpuModel(oracle_lib, "ora", model_file(**GOLDEN_MODEL), 16)
    try:
        toks = G["eval_tokens"]
        n_past = 0
        for i, n in enumerate((4, 1, 1)):
            got = m.eval(nth, n_past, toks[n_past:n_past + n])
            assert np.array_equal(bits(got), bits(G[f"eval_logits_nth{nth}_{i}"]))
            n_past += n
        assert np.array_equal(bits(m.kv(0, 0, n_past)[:, :256]), bits(G[f"eval_k0_nth{nth}"]))
        assert np.array_equal(bits(m.kv(0, 1, n_past)[:, :256]), bits(G[f"eval_v0_nth{nth}"]))
    finally:
        m.free()
