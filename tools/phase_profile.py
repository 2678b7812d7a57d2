"""Per-phase timeline of the whole-token kernel from its built-in %globaltimer stamps (development aid).
Prints, averaged over layers, how long each segment takes for the median CTA and for the slowest CTA."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_swift_b200 as lsb
from llama_swift_b200 import ggml_format as gf

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=8)
ap.add_argument("--pos", type=int, default=64)
ap.add_argument("--out", default="")
ap.add_argument("--tp", type=int, default=1, help="profile a single-process tensor-parallel group of this many GPUs (rank 0's timeline)")
args = ap.parse_args()
path = f"/tmp/probe-7b-l{args.layers}.bin"
if not os.path.exists(path):
    gf.write_synthetic_model(path, gf.HParams(n_layer=args.layers), seed=0, mode="direct")
m = (lsb.llama_model_load(path, n_ctx=max(128, args.pos + 8)) if args.tp == 1 else
     lsb.llama_model_load_group(path, n_ctx=max(128, args.pos + 8), devices=tuple(range(args.tp))))
lsb.llama_eval(m, 8, 0, np.arange(3, 11, dtype=np.int32))
toks, _, ms = m.decode_device(8, 5, args.pos - 8, n_threads=8)     # fill the cache up to pos
print(f"decode {args.pos - 8} steps: {ms / (args.pos - 8) * 1e3:.1f} us/token")
L = lsb.lib()
L.b200_llama_profile_token.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
cap = 148 * (2 + 26 * args.layers + 12) + 1024
buf = np.zeros(cap, dtype=np.int64)
ncta = C.c_int(0)
for rep in range(3):
    marks = L.b200_llama_profile_token(m._h, 8, 5, args.pos, buf.ctypes.data, cap, C.byref(ncta))
assert marks > 0, marks
t = buf[: ncta.value * marks].reshape(ncta.value, marks).astype(np.float64)
t = (t - t[:, 0].min()) / 1e3   # us
# marks per layer (megakernel.cuh): qkv {prologue = wait for inpL + LayerNorm + quantize, rows, epilogue, local barrier},
# attention, then {prologue = wait for the flagged activation words + quantize, rows, epilogue = flagged stores} x 3
names = ["qkv wait (arrival)", "qkv read+norm+quant", "qkv rows", "qkv epilogue", "barrier (qkv->attn)", "attention",
         "wo wait (arrival)", "wo read+quant", "wo rows", "wo epilogue",
         "w13 wait (arrival)", "w13 read+norm+quant", "w13 rows", "w13 epilogue",
         "w2 wait (arrival)", "w2 read+quant", "w2 rows", "w2 epilogue"]
if os.environ.get("B200_PROF_LN") == "1":     # a -DB200_PROF_LN=1 build: 3 extra marks inside each LayerNorm prologue
    ln = lambda p: [p + " read+sum1", p + " sum2+rsqrt", p + " mul+quant", p + " (end)"]
    names = (["qkv wait (arrival)"] + ln("qkv") + ["qkv rows", "qkv epilogue", "barrier (qkv->attn)", "attention",
             "wo wait (arrival)", "wo read+quant", "wo rows", "wo epilogue", "w13 wait (arrival)"] + ln("w13") +
             ["w13 rows", "w13 epilogue", "w2 wait (arrival)", "w2 read+quant", "w2 rows", "w2 epilogue"])
NM = len(names)
nl = args.layers
print(f"layers {nl} pos {args.pos}: kernel span {t[:, NM * nl + 4].max():.1f} us; "
      f"per layer {(t[:, NM * nl].max() - t[:, 0].min()) / nl:.2f} us")
seg = np.zeros((nl, NM, ncta.value))
for il in range(nl):
    for k in range(NM):
        seg[il, k] = t[:, NM * il + k + 1] - t[:, NM * il + k]
for k, n in enumerate(names):
    med = np.median(seg[1:, k, :])
    mx = np.mean(np.max(seg[1:, k, :], axis=1))
    mn = np.mean(np.min(seg[1:, k, :], axis=1))
    print(f"{n:>22}: median CTA {med:6.2f} us   slowest CTA {mx:6.2f} us   fastest {mn:6.2f} us")
print(f"{'sum of medians':>22}: {sum(np.median(seg[1:, k, :]) for k in range(NM)):.2f} us/layer")
o = NM * nl
print(f"output: wait {np.median(t[:, o + 1] - t[:, o]):.2f}  read+norm+quant {np.median(t[:, o + 2] - t[:, o + 1]):.2f}  rows {np.median(t[:, o + 3] - t[:, o + 2]):.2f}  store {np.median(t[:, o + 4] - t[:, o + 3]):.2f} us (median CTA)")
if os.environ.get("B200_PROF_ATT") == "1":     # a -DB200_PROF_ATT=1 build: stamps inside the attention phase (CTAs that own a unit)
    base = 18 * nl + 8
    segs = ["wait for q/k/v of this token", "K.Q", "soft_max", "V.P chains"]
    raw = buf[: ncta.value * marks].reshape(ncta.value, marks).astype(np.float64)
    for k, n in enumerate(segs):
        d = []
        for il in range(1, nl):
            a0, a1 = raw[:, base + 5 * il + k], raw[:, base + 5 * il + k + 1]
            ok = (a0 > 0) & (a1 > 0)
            d.append((a1[ok] - a0[ok]) / 1e3)
        d = np.concatenate(d)
        print(f"   attention / {n:>28}: median {np.median(d):5.2f} us   p90 {np.percentile(d, 90):5.2f} us   max {d.max():5.2f} us")
if args.out:
    np.save(args.out, t)
