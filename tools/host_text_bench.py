"""Host-side callees of the token loop at the real LLaMA vocabulary size (32000 pieces): the reference's llama_tokenize /
llama_sample_top_p_top_k (oracle/_ref) vs the drop-ins of csrc/host_text.cpp, microseconds per token, same results.
CPU only.  """
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import llama_swift_b200 as lsb
from llama_swift_b200 import ggml_format as gf
from conftest import REF_SO, _bind_model_api

N = 32000
rng = np.random.default_rng(0)
# sentencepiece-like vocabulary: 3 control pieces, 256 single bytes, then random "words" of 2-8 lowercase letters
pieces = [b"", b"", b""] + [bytes([b]) for b in range(1, 256)]
seen = set(pieces)
while len(pieces) < N:
    w = bytes(rng.integers(97, 123, size=rng.integers(2, 9)).astype(np.uint8))
    w = (b" " + w) if rng.random() < 0.5 else w
    if w not in seen:
        seen.add(w)
        pieces.append(w)
path = "/tmp/b200_hosttext_32000.bin"
if not os.path.exists(path):
    gf.write_synthetic_model(path + ".tmp", gf.HParams(n_vocab=N, n_layer=1), seed=1, mode="direct", vocab=pieces)
    os.replace(path + ".tmp", path)
L = C.CDLL(REF_SO)
_bind_model_api(L, "ref_llama")
vp, ci, cp = C.c_void_p, C.c_int, C.c_char_p
L.ref_tokenize.argtypes, L.ref_tokenize.restype = [vp, cp, ci, vp, ci], ci
L.ref_sampler_new.argtypes, L.ref_sampler_new.restype = [ci], vp
L.ref_sample_top_p_top_k.argtypes = [vp, vp, vp, vp, ci, C.c_double, ci, C.c_double, C.c_double]
L.ref_sample_top_p_top_k.restype = ci
err = C.create_string_buffer(256)
h = C.c_void_p(L.ref_llama_load(path.encode(), 8, err, 256))
assert h, err.value

text = b"".join(pieces[i] for i in rng.integers(259, N, size=400))
buf = np.empty(len(text) + 2, np.int32)
t0 = time.perf_counter(); n_ref = L.ref_tokenize(h, text, 1, buf.ctypes.data, len(buf)); t_ref = time.perf_counter() - t0
want = buf[:n_ref].copy()
tok = lsb.Tokenizer(pieces=pieces)
t0 = time.perf_counter()
for _ in range(20):
    got = tok(text, bos=True)
t_new = (time.perf_counter() - t0) / 20
assert np.array_equal(got, want)
print(f"tokenizer, {len(text)} bytes -> {n_ref} tokens: reference {t_ref / n_ref * 1e6:.0f} us/token, trie {t_new / n_ref * 1e6:.2f} us/token (identical ids)")

logits = (rng.standard_normal(N) * 4.0).astype(np.float32)
last = rng.integers(0, N, size=64).astype(np.int32)
rs, ms = C.c_void_p(L.ref_sampler_new(-1)), lsb.Sampler(-1)
ids_r, ids_m = [], []
t0 = time.perf_counter()
for _ in range(20):
    ids_r.append(L.ref_sample_top_p_top_k(h, rs, logits.ctypes.data, last.ctypes.data, 64, 1.3, 40, 0.95, 0.8))
t_r = (time.perf_counter() - t0) / 20
t0 = time.perf_counter()
for _ in range(20):
    ids_m.append(ms.sample(logits, last, 1.3, 40, 0.95, 0.8))
t_m = (time.perf_counter() - t0) / 20
assert ids_r == ids_m
print(f"sampler, n_vocab {N}: reference {t_r * 1e6:.0f} us/token, bitmap {t_m * 1e6:.0f} us/token (identical ids)")
