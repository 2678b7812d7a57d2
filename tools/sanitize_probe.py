"""Tiny workload for compute-sanitizer (memcheck / racecheck / initcheck): a 1-layer 7B-width model through every path --
prompt batch on the tensor cores and on the CUDA-core columns, token-by-token llama_eval (whole-token kernel and per-matrix
kernels), and the device-resident greedy loop."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_swift_b200 as lsb
from llama_swift_b200 import ggml_format as gf
path = "/tmp/b200_sanitize_model.bin"
if not os.path.exists(path):
    gf.write_synthetic_model(path, gf.HParams(n_vocab=256, n_layer=1), seed=5)
m = lsb.llama_model_load(path, n_ctx=48)
toks = np.array([1, 7, 9, 11, 13, 15, 17, 19, 21], np.int32)
ref = lsb.llama_eval(m, 8, 0, toks)
m.set_option("tc", 0); a = lsb.llama_eval(m, 8, 0, toks)
m.set_option("batch", 0); b = lsb.llama_eval(m, 8, 0, toks)
m.set_option("mega", 0); c = lsb.llama_eval(m, 8, 0, toks[:3])
m.set_option("mega", 1)
d = lsb.llama_eval(m, 8, 0, toks[:3])
t, _, _ = m.decode_device(9, int(ref.argmax()), 4, n_threads=8)
same = lambda x, y: np.array_equal(x.view(np.uint32), y.view(np.uint32))
print("sanitize probe:", same(ref, a), same(ref, b), same(c, d), t.tolist())
# round 2 additions: the tensor-core mat-mul on a 9-token batch, the sampler's candidate stage behind an evaluation and alone,
# the Q4_1 mat-vec in both of its modes (term / chain split: 28 rows per CTA; one thread per row: 84 rows per CTA)
m.set_option("batch", 1); m.set_option("tc", 1); m.set_option("tc_min_n", 2)
e = lsb.llama_eval(m, 8, 0, toks)
last = np.zeros(64, np.int32)
cand = lsb.llama_eval_topk(m, 8, 9, toks[:1], last)
rng = np.random.default_rng(0)
lg = (rng.standard_normal(32000) * 4).astype(np.float32)
cand2 = lsb.sample_topk(lg, last)
ok41 = []
for M, K in ((4096, 128), (12288, 128)):
    w = gf.quantize_q4_1((rng.standard_normal((M, K)) / np.sqrt(K)).astype(np.float32))
    x = rng.standard_normal(K).astype(np.float32)
    ok41.append(bool(np.isfinite(lsb.q4_1_matvec(w, x)).all()))
print("sanitize probe 2:", same(ref, e), cand is not None, cand2 is not None, ok41)
m.free()
