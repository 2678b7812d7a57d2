"""Prompt-batch probe (development aid): one llama_eval of N tokens on an L-layer 7B-width model; used under ncu."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_swift_b200 as lsb
from llama_swift_b200 import ggml_format as gf
ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=2)
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--tc", type=int, default=1)
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
path = f"/tmp/probe-7b-l{args.layers}.bin"
if not os.path.exists(path):
    gf.write_synthetic_model(path, gf.HParams(n_layer=args.layers), seed=0, mode="direct")
m = lsb.llama_model_load(path, n_ctx=args.n + 8)
m.set_option("tc", args.tc)
toks = np.random.default_rng(0).integers(3, 32000, size=args.n).astype(np.int32)
for r in range(args.reps):
    t0 = time.perf_counter(); lsb.llama_eval(m, 8, 0, toks); dt = time.perf_counter() - t0
    print(f"N={args.n} layers={args.layers} tc={args.tc}: {dt*1e3:.2f} ms, launches {m.last_launches}", flush=True)
