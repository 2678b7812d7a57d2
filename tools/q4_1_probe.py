"""Decode speed of a Q4_1 model at LLaMA-7B width (development aid): us per token at two depths -> us per layer, and the
32-layer extrapolation.  Q4_1 runs through the per-matrix kernels (kernels_q4_1.cuh), one launch per mat-vec."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_swift_b200 as lsb
from llama_swift_b200 import ggml_format as gf

res = {}
for layers in (2, 6):
    path = f"/tmp/probe-7b-q41-l{layers}.bin"
    if not os.path.exists(path):
        gf.write_synthetic_model(path, gf.HParams(n_layer=layers, ftype=gf.FTYPE_Q4_1), seed=0, mode="quantize")
    m = lsb.llama_model_load(path, n_ctx=128)
    tok = np.array([5], np.int32)
    for i in range(8):
        lsb.llama_eval(m, 8, i, tok)
    t0 = time.perf_counter()
    n = 64
    for i in range(n):
        lsb.llama_eval(m, 8, 8 + i, tok)
    dt = (time.perf_counter() - t0) / n
    res[layers] = dt
    print(f"Q4_1 7B width, {layers} layers: {dt * 1e6:.1f} us/token (host loop, logits D2H included), launches {m.last_launches}", flush=True)
    m.free()
per_layer = (res[6] - res[2]) / 4
fixed = res[2] - 2 * per_layer
print(f"per layer {per_layer * 1e6:.1f} us, embedding + output + host {fixed * 1e6:.1f} us -> 32 layers: {(fixed + 32 * per_layer) * 1e6:.0f} us/token = {1 / (fixed + 32 * per_layer):.0f} tok/s")
