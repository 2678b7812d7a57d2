"""Quick GPU probe (development aid): kernel-level mat-vec timings and a short decode-loop timing."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_swift_b200 as lsb
from llama_swift_b200 import ggml_format as gf

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=32)
ap.add_argument("--steps", type=int, default=64)
ap.add_argument("--n-past", type=int, default=8)
ap.add_argument("--matvec", action="store_true")
ap.add_argument("--pdl", type=int, default=0)
ap.add_argument("--mega", type=int, default=1)
args = ap.parse_args()

if args.matvec:
    rng = np.random.default_rng(0)
    for (M, K) in [(4096, 4096), (4096, 11008), (12288, 4096), (22016, 4096), (32000, 4096)]:
        blk = gf._direct_q4_0(rng, M, K, 0.02)
        x = rng.standard_normal(K).astype(np.float32)
        for lp in (1, 2, 4):
            try:
                out, ms = lsb.q4_0_matvec(blk, x, lane_pairs=lp, timed=True)
                gb = M * K / 32 * 20 / 1e9
                print(f"matvec {M}x{K} lp={lp}: {ms*1e3:.1f} us  {gb/ms*1e3:.0f} GB/s", flush=True)
            except Exception as e:
                print(f"matvec {M}x{K} lp={lp}: {e}")

path = f"/tmp/probe-7b-l{args.layers}.bin"
t = time.time()
if not os.path.exists(path):
    gf.write_synthetic_model(path, gf.HParams(n_layer=args.layers), seed=0, mode="direct")
print(f"model file: {time.time()-t:.1f}s", flush=True)
t = time.time()
m = lsb.llama_model_load(path, n_ctx=args.n_past + args.steps + 8)
print(f"load: {time.time()-t:.1f}s, weights {m.weight_bytes/1e9:.3f} GB", flush=True)
m.set_option("pdl", args.pdl)
m.set_option("mega", args.mega)
lsb.llama_eval(m, 8, 0, np.arange(3, 3 + args.n_past, dtype=np.int32))
for rep in range(3):
    toks, _, ms = m.decode_device(args.n_past, 5, args.steps, n_threads=8)
    print(f"decode {args.steps} steps: {ms:.2f} ms -> {args.steps/ms*1e3:.1f} tok/s, {ms/args.steps*1e3:.1f} us/tok, "
          f"W-only {m.weight_bytes/ (ms/args.steps*1e-3)/1e9:.0f} GB/s  launches {m.last_launches}", flush=True)
