#!/usr/bin/env python3
"""Generate tests/golden/reference_vectors.npz from the UNMODIFIED reference (oracle/_ref/libllama_ref.so, built by
oracle/Makefile from /root/reference).  The reference ships no golden vectors of its own, so
these are outputs of the reference itself, run in the authoring container with the pinned flags
(-O3 -DNDEBUG -std=c11 -mavx -mavx2 -mfma -mf16c -msse3), 1 thread unless stated.  Everything is seeded; the model
file behind the llama_eval vectors is re-created from its seed by the test (llama.swift_b200/ggml_format.py).

usage: python tests/golden/make_golden.py      (needs /root/reference; re-run only when the recipe changes)
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import CpuModel, _bind_model_api, model_file  # noqa: E402

GOLDEN_MODEL = dict(n_layer=1, n_vocab=256, seed=11)


def main():
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libllama_ref.so"))
    _bind_model_api(L, "ref_llama")
    vp, ci = C.c_void_p, C.c_int
    L.ref_mul_mat_q4.argtypes = [ci, vp, ci, ci, vp, ci, vp, ci]
    L.ref_unary_op.argtypes = [ci, vp, ci, ci, ci, ci, vp, ci]
    L.ref_quantize_row.argtypes = [ci, vp, vp, ci]
    L.ref_quantize_weights.argtypes = [ci, vp, vp, ci, ci]
    L.ref_quantize_weights.restype = C.c_size_t
    rng = np.random.default_rng(20230313)
    out = {}

    x = (rng.standard_normal(256) * 1.7).astype(np.float32)
    x[:32] = 0
    x[32] = 7.0; x[33:48] = np.arange(-7, 8, dtype=np.float32)[:15] + 0.5      # ties: nearest-even on AVX2
    out["q_x"] = x
    for t, nbytes in ((2, 160), (3, 192)):
        q = np.zeros(nbytes, np.uint8)
        L.ref_quantize_row(t, x.ctypes.data, q.ctypes.data, 256)
        out[f"q_row_type{t}"] = q

    M, K, N = 48, 256, 2
    w = (rng.standard_normal((M, K)) / 16).astype(np.float32)
    xx = (rng.standard_normal(N * K) * 1.3).astype(np.float32)
    out["mm_w"], out["mm_x"] = w, xx
    for t, bpb in ((2, 20), (3, 24)):
        wq = np.zeros(M * K // 32 * bpb, np.uint8)
        L.ref_quantize_weights(t, w.copy().ctypes.data, wq.ctypes.data, M * K, K)
        y = np.zeros(N * M, np.float32)
        L.ref_mul_mat_q4(t, wq.ctypes.data, M, K, xx.ctypes.data, N, y.ctypes.data, 1)
        out[f"mm_wq_type{t}"], out[f"mm_y_type{t}"] = wq, y

    v = (rng.standard_normal(512) * 2.0).astype(np.float32)
    out["u_x"] = v
    for name, op, shape, n_past in (("norm", 0, (512, 1, 1), 0), ("silu", 1, (512, 1, 1), 0), ("soft_max", 2, (512, 1, 1), 0),
                                    ("rope_p5", 3, (128, 4, 1), 5)):
        y = np.zeros(512, np.float32)
        L.ref_unary_op(op, v.ctypes.data, shape[0], shape[1], shape[2], n_past, y.ctypes.data, 1)
        out["u_" + name] = y

    path = model_file(**GOLDEN_MODEL)
    toks = [np.array([1, 17, 200, 33], np.int32), np.array([5], np.int32), np.array([77], np.int32)]
    for nth in (1, 8):
        m = CpuModel(L, "ref_llama", path, 16)
        n_past = 0
        for i, t in enumerate(toks):
            out[f"eval_logits_nth{nth}_{i}"] = m.eval(nth, n_past, t)
            n_past += len(t)
        out[f"eval_k0_nth{nth}"] = m.kv(0, 0, n_past)[:, :256].copy()
        out[f"eval_v0_nth{nth}"] = m.kv(0, 1, n_past)[:, :256].copy()
        m.free()
    out["eval_tokens"] = np.concatenate(toks)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
