"""SURVEY.md section 7.1(a) noise-floor gate for a model file: the UNMODIFIED reference (oracle/_ref) run with 1 thread
against the same reference run with 8 threads over one teacher-forced token stream.  The only difference between the two
runs is the summation order of the V*P partial sums (ggml.c:5553-5577), i.e. exactly the kind of noise a kernel that
mirrors the AVX2 arithmetic may have; a model file passes when the p99 rel-L2 between the two is below 1e-4 (chaotic
files -- plain random weights -- fail by orders of magnitude).  CPU only."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=32)
ap.add_argument("--steps", type=int, default=24)
ap.add_argument("--out", default="")
args = ap.parse_args()

path = bench.ensure_model(args.layers)
L = bench.ref_lib()
assert L is not None, "oracle/_ref/libllama_ref.so missing (make -C oracle ref)"
err = C.create_string_buffer(512)


def run(n_threads, stream):
    h = C.c_void_p(L.ref_llama_load(path.encode(), 64, err, 512))
    assert h, err.value
    out = []
    logits = np.empty(32000, np.float32)
    p = np.array(bench.PROMPT, np.int32)
    L.ref_llama_eval(h, n_threads, 0, p.ctypes.data, len(p), logits.ctypes.data, err, 512)
    out.append(logits.copy())
    for i, t in enumerate(stream):
        tt = np.array([t], np.int32)
        L.ref_llama_eval(h, n_threads, len(p) + i, tt.ctypes.data, 1, logits.ctypes.data, err, 512)
        out.append(logits.copy())
    L.ref_llama_free(h)
    return np.stack(out)


# the stream = the 8-thread run's own greedy tokens (what bench.py generates), then replayed with 1 thread
h = C.c_void_p(L.ref_llama_load(path.encode(), 64, err, 512))
logits = np.empty(32000, np.float32)
p = np.array(bench.PROMPT, np.int32)
L.ref_llama_eval(h, 8, 0, p.ctypes.data, len(p), logits.ctypes.data, err, 512)
stream = []
cur = int(logits.argmax())
for i in range(args.steps):
    stream.append(cur)
    tt = np.array([cur], np.int32)
    L.ref_llama_eval(h, 8, len(p) + i, tt.ctypes.data, 1, logits.ctypes.data, err, 512)
    cur = int(logits.argmax())
L.ref_llama_free(h)

a, b = run(8, stream), run(1, stream)
rel = np.linalg.norm(a.astype(np.float64) - b, axis=1) / np.linalg.norm(a.astype(np.float64), axis=1)
top2 = np.sort(a, axis=1)[:, -2:]
res = {"model": os.path.basename(path), "layers": args.layers, "evals": int(len(rel)), "comparison": "oracle/_ref 8 threads vs 1 thread, same teacher-forced stream",
       "rel_l2_median": float(np.median(rel)), "rel_l2_p99": float(np.percentile(rel, 99)), "rel_l2_max": float(rel.max()),
       "argmax_agree": float(np.mean(a.argmax(1) == b.argmax(1))), "bit_identical_evals": int(sum(np.array_equal(a[i].view(np.uint32), b[i].view(np.uint32)) for i in range(len(rel)))),
       "top2_margin_median": float(np.median(top2[:, 1] - top2[:, 0])), "gate": "p99 < 1e-4", "pass": bool(np.percentile(rel, 99) < 1e-4)}
print(json.dumps(res, indent=1))
if args.out:
    json.dump(res, open(args.out, "w"), indent=1)
