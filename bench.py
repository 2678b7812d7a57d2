#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric: decode tokens/sec, LLaMA-7B Q4_0, bs = 1, and % of the HBM roofline.

One "step" = one pass of the hot path = llama_eval of ONE token (PO.mm:510-735) at the next position of a
512-token generation that follows an 8-token prompt (BASELINE.json configs[1]).  Synthetic model: a 7B-shaped file in
the reference's own "ggml" format (seeded Q4_0 blocks, llama.swift_b200/ggml_format.py) -- no real weights exist here.

  value ............ tokens/s of the device-resident loop (weights, KV cache, token id all in HBM when the clock
                     starts; greedy arg-max on the GPU feeds the next step), CUDA-event timed on the launching stream
  e2e .............. the same generation driven through the reference-facing C ABI (b200_llama_eval, the llama_eval
                     drop-in): per step the token id goes host->device and the 32000 logits come back device->host
                     (pinned staging inside the library) and the host picks the arg-max
  roofline ......... dominant kernel = decode_token_kernel (the whole token in one persistent launch): algorithmic
                     bytes per launch (weights once + norm weights + embedding row + f32 KV
                     read/write) / its average launch duration from CUDA events around each launch
  cpu_baseline ..... the reference's own CPU path (oracle/_ref = unmodified ggml.c + llama_eval) on this host's cores

--impl reference times ONLY that CPU path (rank 0; other ranks exit) and prints the same JSON shape.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PROMPT = [1, 15043, 3186, 29892, 590, 1024, 338, 29871]      # 8 fixed ids, BOS first (utils.cpp:284-286)
N_PROMPT = len(PROMPT)

# algorithmic bytes (BASELINE.md section 2 / SURVEY.md section 8d), Q4_0: weights once (20 B per 32) + f32 norm weights + one
# embedding row + f32 KV rows read / written
MODELS = {"7b": dict(n_embd=4096, n_head=32, n_layer=32, n_ff=11008, n_vocab=32000),          # PO.mm:41-50
          "13b": dict(n_embd=5120, n_head=40, n_layer=40, n_ff=13824, n_vocab=32000)}         # two-part file (PO.mm:35), BASELINE.json configs[3]
MODEL = "7b"


def _geom(layers=None):
    g = dict(MODELS[MODEL])
    if layers:
        g["n_layer"] = layers
    return g


def weight_bytes(layers=None) -> int:
    g = _geom(layers)
    e, f, v, l = g["n_embd"], g["n_ff"], g["n_vocab"], g["n_layer"]
    return (l * (4 * e * e + 3 * e * f) + v * e) // 32 * 20


W_BYTES = 4_129_423_360       # LLaMA-7B Q4_0 (== weight_bytes() for the default model)


def algorithmic_bytes(pos: int, layers=None) -> int:
    g = _geom(layers)
    e, l = g["n_embd"], g["n_layer"]
    s_bytes = (2 * l + 1) * e * 4 + e // 32 * 20
    kv_row = 2 * l * e * 4            # K and V, f32, per cached position
    return weight_bytes(layers) + s_bytes + kv_row * (pos + 1) + kv_row


def model_path(layers: int) -> str:
    d = os.environ.get("B200_BENCH_DIR", "/tmp/b200_bench")
    os.makedirs(d, exist_ok=True)
    full = MODELS[MODEL]["n_layer"]
    tag = "" if MODEL == "7b" else "-" + MODEL
    return os.path.join(d, f"ggml-model{tag}-q4_0.bin" if layers == full else f"ggml-model{tag}-q4_0-l{layers}.bin")


def ensure_model(layers: int) -> str:
    from llama_swift_b200 import ggml_format as gf
    path = model_path(layers)
    n_parts = gf.LLAMA_N_PARTS[MODELS[MODEL]["n_embd"]]
    if not all(os.path.exists(path if p == 0 else f"{path}.{p}") for p in range(n_parts)):
        tmp = path + f".tmp{os.getpid()}"
        g = MODELS[MODEL]
        gf.write_synthetic_model(tmp, gf.HParams(n_vocab=g["n_vocab"], n_embd=g["n_embd"], n_head=g["n_head"], n_layer=layers), seed=0, mode="direct")
        for p in range(n_parts):
            os.replace(tmp + ("" if p == 0 else f".{p}"), path if p == 0 else f"{path}.{p}")
    return path


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md recipe).  NVML in-process (a query costs
    ~1 ms, so even a 0.2 s region gets dozens of samples); falls back to polling `nvidia-smi` when pynvml is missing."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    MASKS = [0x8, 0x40, 0x20, 0x4]       # nvmlClocksThrottleReason{HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.source = index, [], False, "nvml"
        self.nv = self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv, self.source = None, "nvidia-smi"

    def _nvml_sample(self):
        nv, h = self.nv, self.h
        sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
        try:
            reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
        except Exception:
            reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
        try:
            power = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
        except Exception:
            power = 0.0
        return [sm, self.sm_max, power] + [bool(reasons & m) for m in self.MASKS]

    def _smi_sample(self):
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if not out:
            return None
        f = [x.strip() for x in out.split(",")]
        return [float(f[0]), float(f[1]), float(f[2])] + [x.lower().startswith("active") for x in f[3:7]]

    def run(self):
        while not self.stop_flag:
            try:
                smp = self._nvml_sample() if self.nv is not None else self._smi_sample()
                if smp is not None:
                    self.samples.append(smp)
            except Exception:
                pass
            time.sleep(0.005 if self.nv is not None else 0.05)

    def summary(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples (neither NVML nor nvidia-smi answered)"]}
        sm = sorted(s[0] for s in self.samples)
        reasons = [n for i, n in enumerate(self.NAMES) if any(s[3 + i] for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1], "reasons": reasons,
                "power_w_max": max(s[2] for s in self.samples), "samples": len(self.samples), "source": self.source}


def ref_lib():
    so = os.path.join(ROOT, "oracle", "_ref", "libllama_ref.so")
    if not os.path.exists(so):
        if os.path.isdir("/root/reference/Sources/cpp"):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
        else:
            return None
    L = C.CDLL(so)
    L.ref_llama_load.restype = C.c_void_p
    L.ref_llama_load.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
    L.ref_llama_eval.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_char_p, C.c_size_t]
    L.ref_llama_free.argtypes = [C.c_void_p]
    return L


def time_reference_cpu(path: str, n_threads: int, budget_s: float, max_steps: int, keep=None, n_vocab: int = 32000):
    """The reference's own llama_eval on the host: 8-token prompt, then greedy single-token steps.  Returns
    (tokens/s over the timed decode steps, steps timed, description).  keep: optional dict that receives the run's
    outputs -- "first" (arg-max after the prompt), "tokens" (arg-max of every step) and "logits" [steps, n_vocab] -- for
    the parity check against the GPU stream."""
    L = ref_lib()
    if L is None:
        return None, 0, "oracle/_ref/libllama_ref.so not available"
    err = C.create_string_buffer(512)
    h = L.ref_llama_load(path.encode(), 64, err, 512)
    if not h:
        return None, 0, "reference load failed: " + err.value.decode()
    h = C.c_void_p(h)
    logits = np.empty(n_vocab, np.float32)
    toks = np.array(PROMPT, np.int32)
    L.ref_llama_eval(h, n_threads, 0, toks.ctypes.data, len(toks), logits.ctypes.data, err, 512)   # prompt, untimed
    n_past, cur = len(toks), int(logits.argmax())
    if keep is not None:
        keep.update({"first": cur, "tokens": [], "logits": []})
    times = []
    t_begin = time.perf_counter()
    while len(times) < max_steps and n_past < 62:
        t = np.array([cur], np.int32)
        t0 = time.perf_counter()
        L.ref_llama_eval(h, n_threads, n_past, t.ctypes.data, 1, logits.ctypes.data, err, 512)
        times.append(time.perf_counter() - t0)          # the reference's own t_predict_us window (PO.mm:837-845)
        n_past += 1
        cur = int(logits.argmax())
        if keep is not None:
            keep["tokens"].append(cur)
            keep["logits"].append(logits.copy())
        if time.perf_counter() - t_begin > budget_s and len(times) >= 3:
            break
    L.ref_llama_free(h)
    total = sum(times)
    med = sorted(times)[len(times) // 2]
    return len(times) / total, len(times), (f"reference ggml CPU path (oracle/_ref, AVX2 build), {n_threads} threads, 8-token prompt then "
                                             f"{len(times)} greedy decode steps at n_past 8..{n_past - 1}; median {med * 1e3:.1f} ms/token")


def calibrated_reference_cpu(path: str, budget_s: float, max_steps: int):
    """The reference CPU arm with "all the host threads it can use": its spin-wait pool (ggml.c:9061-9107, re-created on
    every llama_eval) gets SLOWER past a point, so every candidate thread count is tried on 3 steps and the fastest one is
    timed.  Used by BOTH the cpu_baseline leg and --impl reference, so the two report the same thread count."""
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64) if c <= cores} | {min(cores, 8)})
    cal = {}
    for c in cands:
        r, _, _ = time_reference_cpu(path, c, budget_s=20.0, max_steps=3)
        if r is not None:
            cal[c] = r
    nth = max(cal, key=cal.get) if cal else min(cores, 8)
    tps, n, desc = time_reference_cpu(path, nth, budget_s=budget_s, max_steps=max_steps)
    if tps is not None:
        desc += ("; thread-count calibration tok/s: " + ", ".join(f"{c}: {v:.2f}" for c, v in sorted(cal.items())) +
                 f" of {cores} host cores (8 = the Swift default, LlamaRunner.swift:17)")
    return tps, n, desc, nth, cores


PARITY_STEPS = 24


def parity_reference(path: str, n_threads: int, n_vocab: int):
    """Rank 0: PARITY_STEPS greedy steps of the UNMODIFIED reference (oracle/_ref) after the 8-token prompt, with the thread
    count the GPU run mirrors (the reference's V*P summation order depends on it).  Returns the kept outputs or None."""
    keep = {}
    tps, n, _ = time_reference_cpu(path, n_threads, budget_s=120.0, max_steps=PARITY_STEPS, keep=keep, n_vocab=n_vocab)
    if tps is None or n < 1:
        return None
    return keep


def parity_block(model, lsb, ref, n_threads: int):
    """The GPU stream over the same steps, teacher-forced with the reference's tokens so that one flipped arg-max cannot
    hide the steps after it; every rank of a group makes the same calls, rank 0 compares."""
    toks = np.array(ref["tokens"], np.int32)
    n = len(toks)
    got_first = int(lsb.llama_eval(model, n_threads, 0, np.array(PROMPT, np.int32)).argmax())
    forced = toks                                            # token fed at step i+1 = the reference's arg-max of step i
    gtoks, glogits, _ = model.decode_device(N_PROMPT, ref["first"], n, n_threads=n_threads, forced_tokens=forced, want_logits=True)
    want = np.stack(ref["logits"]).astype(np.float64)
    got = glogits.astype(np.float64)
    rel = np.linalg.norm(got - want, axis=1) / np.maximum(np.linalg.norm(want, axis=1), 1e-30)
    bit = int(sum(np.array_equal(glogits[i].view(np.uint32), ref["logits"][i].view(np.uint32)) for i in range(n)))
    return {"steps": n, "positions": [N_PROMPT, N_PROMPT + n - 1], "vs": "oracle/_ref (unmodified reference llama_eval), %d threads" % n_threads,
            "argmax_equal": bool(got_first == ref["first"] and np.array_equal(gtoks, toks)), "max_rel_l2": float(rel.max()),
            "bit_identical_steps": bit, "tolerance": 1e-3, "ok": bool(got_first == ref["first"] and np.array_equal(gtoks, toks) and rel.max() <= 1e-3)}



# ---- configs[2]: 2048-token prefill through the batch path -----------------------------------------------------------------------
PREFILL_FLOP_PER_TOKEN = 2 * 6_607_077_376          # the 225 mat-muls; lm_head counted like the reference (all rows)


def prefill_tokens(n):
    return np.random.default_rng(1234).integers(3, 32000, size=n).astype(np.int32)


def bench_prefill(args, steps, warmup):
    import torch
    import llama_swift_b200 as lsb
    n = args.prompt_tokens
    path = ensure_model(args.layers)
    model = lsb.llama_model_load(path, n_ctx=n + 8, device=0)
    toks = prefill_tokens(n)
    # parity at a size the CPU reference evaluates in seconds: the first 64 tokens as ONE batched llama_eval call
    parity = None
    if not args.no_parity:
        L = ref_lib()
        if L is not None:
            err = C.create_string_buffer(512)
            h = L.ref_llama_load(path.encode(), 72, err, 512)
            if h:
                h = C.c_void_p(h)
                want = np.empty(model.n_vocab, np.float32)
                t64 = np.ascontiguousarray(toks[:64])
                # the reference sizes its scratch buffer from the 4-token probe call it always makes first (PO.mm:822, 527-541)
                probe = np.array([0, 1, 2, 3], np.int32)
                L.ref_llama_eval(h, args.threads, 0, probe.ctypes.data, 4, want.ctypes.data, err, 512)
                t0 = time.perf_counter()
                L.ref_llama_eval(h, args.threads, 0, t64.ctypes.data, 64, want.ctypes.data, err, 512)
                cpu_s = time.perf_counter() - t0
                L.ref_llama_free(h)
                got = lsb.llama_eval(model, args.threads, 0, t64)
                rel = float(np.linalg.norm(got.astype(np.float64) - want) / max(np.linalg.norm(want.astype(np.float64)), 1e-30))
                parity = {"prompt_tokens": 64, "vs": "oracle/_ref (unmodified reference llama_eval, one N = 64 call), %d threads" % args.threads,
                          "argmax_equal": bool(int(got.argmax()) == int(want.argmax())), "max_rel_l2": rel,
                          "bit_identical": bool(np.array_equal(got.view(np.uint32), want.view(np.uint32))), "tolerance": 1e-3,
                          "ok": bool(rel <= 1e-3 and int(got.argmax()) == int(want.argmax())), "reference_cpu_tokens_per_s": 64 / cpu_s}
    for _ in range(warmup):
        lsb.llama_eval(model, args.threads, 0, toks)
    sampler = ClockSampler(0)
    sampler.start()
    torch.cuda.synchronize()
    dev_ms, wall = [], []
    for _ in range(steps):
        t0 = time.perf_counter()
        lsb.llama_eval(model, args.threads, 0, toks)       # host token ids in, last token's logits out (pinned staging in the library)
        wall.append(time.perf_counter() - t0)
        dev_ms.append(model.last_eval_ms)
    torch.cuda.synchronize()
    clocks = sampler.summary()
    launches = model.last_launches
    ms = float(np.mean(dev_ms))
    tps = n / (ms * 1e-3)
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = (float(json.load(open(peaks_file))["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)") \
        if os.path.exists(peaks_file) else (1400.0, "fallback (B200_PROFILING.md)")
    flops = PREFILL_FLOP_PER_TOKEN * n + 2.0 * n * n * 4096 * 32        # + causal attention ~ 2 N^2 n_embd n_layer
    achieved = flops / (ms * 1e-3) / 1e12
    tc_prof = None
    tpf = os.path.join(ROOT, "profiles", "r2_prefill_tc_ncu.json")
    if os.path.exists(tpf):
        tc_prof = json.load(open(tpf))
    line = {"metric": "prefill tokens/sec LLaMA-7B Q4_0, %d-token prompt" % n, "value": tps, "unit": "tokens/s", "n_gpus": 1, "steps": steps,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int4 x int4 block dots as exact fp16 x fp16 -> f32 tcgen05.mma (TMEM accumulators), fp32 lane accumulation (bit-exact AVX2 order), f32 KV",
            "data": "synthetic",
            "config": {"workload": "LLaMA-7B Q4_0 prefill, %d-token prompt in one llama_eval call (BASELINE.json configs[2])" % n,
                       "model_file": "synthetic ggml-format 7B (n_embd 4096, n_layer %d, n_vocab 32000), seed 0" % args.layers,
                       "n_ctx": n + 8, "chunking": "256-token chunks inside the library", "ref_threads_mirrored": args.threads,
                       "l2_policy": "per-step working set (4.13 GB weights + activations) >> 126 MB L2, no flush needed", "parallelism": "1 GPU"},
            "clocks": clocks,
            "e2e": {"value": n / float(np.mean(wall)), "unit": "tokens/s", "h2d_bytes_per_step": int(n * 4), "d2h_bytes_per_step": int(model.n_vocab * 4),
                    "note": "b200_llama_eval wall time: token ids from host memory, logits back to host memory"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "q4_gemm_tc_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": None,
                         "note": "algorithmic flops = 2 x 6.607e9 per token + causal attention; the exact per-lane fp32 fma chain of Q4_0 x Q4_0 needs the "
                                 "accumulators out of TMEM after every 32-element block, and what the SM must issue for that drain and for the nibble unpack "
                                 "bounds the kernel (DESIGN.md section 4.4), not the tensor pipe",
                         "ncu": tc_prof},
            "cpu_baseline": None if parity is None else {"value": parity["reference_cpu_tokens_per_s"], "unit": "tokens/s", "cores": args.threads, "kind": "reference",
                                                          "sample": "reference llama_eval of the first 64 prompt tokens in one call (N = 64), %d threads" % args.threads},
            "parity": parity}
    print(json.dumps(line))
    model.free()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=512)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="decode", choices=["decode", "prefill"],
                    help="decode = BASELINE.json's metric (configs[1]); prefill = configs[2]: one llama_eval of a 2048-token prompt "
                         "(batch path: tcgen05 / TMEM mat-mul), a step = one whole prompt")
    ap.add_argument("--prompt-tokens", type=int, default=2048)
    ap.add_argument("--model", default="7b", choices=sorted(MODELS), help="7b = BASELINE.json's metric; 13b = configs[3] (two-part file)")
    ap.add_argument("--layers", type=int, default=0, help="debug: fewer layers (the result is then NOT the benchmark)")
    ap.add_argument("--threads", type=int, default=8, help="reference thread count mirrored by the V*P partition (Swift default 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the 24-step comparison with the reference CPU path")
    ap.add_argument("--no-replicas", action="store_true", help="N > 1: skip the informational run of N independent generations")
    ap.add_argument("--parallelism", default="tp", choices=["tp", "replicas"],
                    help="N > 1: tp = ONE bs=1 generation over all GPUs (rows of every matrix split over the ranks; the metric "
                         "BASELINE.json names); replicas = N independent generations")
    args = ap.parse_args()
    global MODEL
    MODEL = args.model
    full_layers = MODELS[MODEL]["n_layer"]
    if not args.layers:
        args.layers = full_layers
    mname = "LLaMA-" + MODEL.upper()[:-1] + "B"

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = args.gpus
    steps, warmup = args.steps, max(3, args.warmup)

    config = {"workload": "%s Q4_0 bs=1 decode, %d-token gen after an 8-token prompt%s" %
                          (mname, steps, (" (BASELINE.json configs[1])" if steps == 512 else " (BASELINE.json configs[1] is the same with 512 tokens)") if MODEL == "7b"
                           else " (BASELINE.json configs[3])"),
              "model_file": "synthetic ggml-format %s (n_embd %d, n_layer %d, n_vocab 32000), seed 0" % (MODEL.upper(), MODELS[MODEL]["n_embd"], args.layers),
              "n_past_start": N_PROMPT, "positions": [N_PROMPT, N_PROMPT + steps - 1], "ref_threads_mirrored": args.threads,
              "l2_policy": "per-step working set (%.2f GB weights) >> 126 MB L2, no flush needed" % (weight_bytes() / 1e9)}
    if args.layers != full_layers:
        config["workload"] += f" [DEBUG: {args.layers} layers -- not the benchmark]"

    if args.impl == "reference" and args.mode == "prefill":
        if rank != 0:
            return 0
        # bounded sample of the prefill workload: the first 64 prompt tokens as ONE batched reference llama_eval per step
        path = ensure_model(args.layers)
        L = ref_lib()
        if L is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libllama_ref.so not available"}))
            return 0
        cores = os.cpu_count() or 1
        nth = min(cores, 16)
        err = C.create_string_buffer(512)
        h = C.c_void_p(L.ref_llama_load(path.encode(), 72, err, 512))
        t64 = np.ascontiguousarray(prefill_tokens(args.prompt_tokens)[:64])
        out = np.empty(32000, np.float32)
        probe = np.array([0, 1, 2, 3], np.int32)      # sizes the reference's scratch buffer (PO.mm:822, 527-541)
        L.ref_llama_eval(h, nth, 0, probe.ctypes.data, 4, out.ctypes.data, err, 512)
        times = []
        for i in range(1 + max(1, min(steps, 5))):
            t0 = time.perf_counter()
            L.ref_llama_eval(h, nth, 0, t64.ctypes.data, 64, out.ctypes.data, err, 512)
            if i > 0:
                times.append(time.perf_counter() - t0)
        L.ref_llama_free(h)
        tps = 64 / float(np.mean(times))
        desc = "reference ggml CPU path (oracle/_ref), %d threads of %d host cores: llama_eval of a 64-token slice of the prompt in one call" % (nth, cores)
        print(json.dumps({"impl": "reference", "metric": "prefill tokens/sec LLaMA-7B Q4_0, %d-token prompt" % args.prompt_tokens, "value": tps,
                          "unit": "tokens/s", "n_gpus": n_gpus, "steps": len(times), "warmup": 1, "ms_per_step": 1e3 * float(np.mean(times)),
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int4 x int4 -> int32 block dots, fp32 accumulate (AVX2 CPU)",
                          "data": "synthetic", "config": {"workload": "LLaMA-7B Q4_0 prefill (BASELINE.json configs[2]); CPU sample: 64 tokens per step"},
                          "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": nth, "kind": "reference", "sample": desc},
                          "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    if args.impl == "reference":
        if rank != 0:
            return 0
        path = ensure_model(args.layers)
        tps, n, desc, nth, cores = calibrated_reference_cpu(path, budget_s=120.0, max_steps=max(3, min(steps, 54)))
        if tps is None:
            print(json.dumps({"impl": "reference", "unavailable": desc}))
            return 0
        line = {"impl": "reference", "metric": "decode tokens/sec %s Q4_0 bs=1" % mname, "value": tps, "unit": "tokens/s",
                "n_gpus": n_gpus, "steps": n, "warmup": 1, "ms_per_step": 1e3 / tps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "int4 x int4 -> int32 block dots, fp32 accumulate (AVX2 CPU)",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": nth, "kind": "reference", "sample": desc},
                "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    if args.mode == "prefill":
        if rank != 0:
            return 0
        return bench_prefill(args, max(1, min(steps, 5)) if args.steps != 512 else 3, warmup)

    import torch
    import llama_swift_b200 as lsb

    dist = None
    saved_stdout = None
    if world > 1:
        # stdout carries exactly ONE JSON line: whatever the libraries print while the ranks talk to each other (NCCL's version
        # banner, warnings) goes to stderr -- file descriptor 1 is pointed at stderr until the line is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) off it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        ensure_model(args.layers)
    if dist is not None:
        dist.barrier()
    path = model_path(args.layers)

    n_ctx = N_PROMPT + steps + 8
    from llama_swift_b200 import dist_util
    tp = world > 1 and args.parallelism == "tp"
    tp_note = None
    model = None
    if tp:
        # one model over all ranks: this rank's row shard + peer-mapped exchange areas (IPC handles over the process group)
        ok = 1
        try:
            model = dist_util.tp_load(lsb, path, n_ctx, local_rank)
        except Exception as e:  # noqa: BLE001 -- every rank must learn about a failure anywhere
            ok, tp_note = 0, f"{type(e).__name__}: {e}"
        flag = torch.tensor([ok], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if model is not None:
                model.free()
            model, tp = None, False
            tp_note = "tensor-parallel setup failed on some rank (%s); ran independent replicas instead" % (tp_note or "other rank")
    if model is None:
        model = lsb.llama_model_load(path, n_ctx=n_ctx, device=local_rank)
    n_vocab = model.n_vocab

    def prompt():
        return lsb.llama_eval(model, args.threads, 0, np.array(PROMPT, np.int32))

    # ---- parity at the benched configuration: the first PARITY_STEPS steps against the unmodified reference (rank 0 runs
    # the CPU side and shares its tokens; every rank of a group then makes the same GPU calls) ----
    parity = None
    if not args.no_parity:
        ref = parity_reference(path, args.threads, n_vocab) if rank == 0 else None
        if dist is not None:
            box = [ref]
            dist.broadcast_object_list(box, src=0)
            ref = box[0]
        if ref is not None:
            parity = parity_block(model, lsb, ref, args.threads)
        elif rank == 0:
            parity = {"steps": 0, "ok": False, "note": "oracle/_ref/libllama_ref.so not available on this box"}

    # ---- warm-up (untimed): prompt + W decode steps, then rewind to the end of the prompt ----
    first = int(prompt().argmax())
    model.decode_device(N_PROMPT, first, warmup, n_threads=args.threads)
    first = int(prompt().argmax())

    def sync_all():
        dist_util.barrier_and_sync(torch.cuda.synchronize)

    # ---- value: device-resident loop, exactly `steps` steps ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    sync_all()
    toks, _, ms = model.decode_device(N_PROMPT, first, steps, n_threads=args.threads)
    sync_all()
    clocks = sampler.summary()
    launches = model.last_launches
    ms_value = dist_util.max_over_ranks(ms, "cuda")

    # ---- roofline: per-launch duration of the token kernel from CUDA events around every launch ----
    first = int(prompt().argmax())
    model.set_option("time_kernel", 1)
    model.decode_device(N_PROMPT, first, steps, n_threads=args.threads)
    kernel_ms = model.kernel_ms_total
    model.set_option("time_kernel", 0)

    # ---- e2e: the C-ABI llama_eval per token with host buffers ----
    first = int(prompt().argmax())
    cur = first
    tok = np.empty(1, np.int32)
    sync_all()
    t0 = time.perf_counter()
    for i in range(steps):
        tok[0] = cur
        logits = lsb.llama_eval(model, args.threads, N_PROMPT + i, tok)     # H2D token, kernels, D2H logits, sync
        cur = int(logits.argmax())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_s = dist_util.max_over_ranks(e2e_s, "cuda")

    # ---- for information at N > 1: what the same GPUs deliver as N independent bs = 1 generations (weak scaling) ----
    replicas = None
    if tp and not args.no_replicas:
        os.environ["B200_PREFILL_COPY"] = "0"          # decode only: no second weight layout
        rep = lsb.llama_model_load(path, n_ctx=n_ctx, device=local_rank)
        rfirst = int(lsb.llama_eval(rep, args.threads, 0, np.array(PROMPT, np.int32)).argmax())
        rep.decode_device(N_PROMPT, rfirst, warmup, n_threads=args.threads)
        rfirst = int(lsb.llama_eval(rep, args.threads, 0, np.array(PROMPT, np.int32)).argmax())
        sync_all()
        _, _, rms = rep.decode_device(N_PROMPT, rfirst, steps, n_threads=args.threads)
        sync_all()
        rms = dist_util.max_over_ranks(rms, "cuda")
        rep.free()
        replicas = {"value": world * steps / (rms * 1e-3), "unit": "tokens/s", "scaling": "weak",
                    "note": "%d independent bs=1 generations, one whole model per GPU, same step count; device-resident loop, max over ranks" % world}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    jobs = 1 if tp else world            # tp: ONE generation over all GPUs; replicas: one per GPU
    tps = jobs * steps / (ms_value * 1e-3)
    mean_bytes = float(np.mean([algorithmic_bytes(N_PROMPT + i) for i in range(steps)])) if args.layers == full_layers else None
    if mean_bytes is not None and tp:
        mean_bytes /= world              # bytes one GPU streams per launch: its row shard of the weights, its heads' KV
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        peak, peak_src = float(json.load(open(peaks_file))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # DRAM bytes per launch come from an `ncu --set full` capture of this kernel on this workload (1 GPU, 32 layers); there is
    # no capture of a tensor-parallel shard (ncu must not wrap a multi-rank run), so N > 1 reports null
    traffic, traffic_src = None, None
    for tf in ("r2_traffic.json", "r1_traffic.json"):
        tfp = os.path.join(ROOT, "profiles", tf)
        if world == 1 and MODEL == "7b" and args.layers == full_layers and os.path.exists(tfp):
            traffic, traffic_src = json.load(open(tfp)).get("dram_bytes_per_launch"), "profiles/" + tf
            break
    roofline = None
    if mean_bytes is not None and kernel_ms:
        achieved = mean_bytes / (kernel_ms / steps * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "decode_token_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": mean_bytes, "weights_only_GBps": weight_bytes() / (world if tp else 1) / (kernel_ms / steps * 1e-3) / 1e9,
                    "kernel_us_per_launch": kernel_ms / steps * 1e3}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        ctps, n, desc, nth, cores = calibrated_reference_cpu(path, budget_s=25.0, max_steps=24)
        if ctps is not None:
            cpu = {"value": ctps, "unit": "tokens/s", "cores": nth, "kind": "reference", "sample": desc,
                   "host_cores_available": cores}
        else:
            cpu = {"value": None, "unit": "tokens/s", "cores": 0, "kind": "reference", "sample": desc}

    if world == 1:
        par = "1 GPU"
    elif tp:
        par = (f"tp{world}: one bs=1 generation over {world} GPUs, every matrix split by rows (wq/wk/wv by head), activation slices "
               "all-gathered by peer-to-peer stores over NVLink inside the token kernel (no NCCL on the data path); "
               "roofline is per GPU (bytes of its shard)")
    else:
        par = f"{world} independent replicas" + (f" [{tp_note}]" if tp_note else "")
    config.update({"parallelism": par, "n_ctx": n_ctx,
                   "graph": "CUDA graph replay of [memset, decode_token_kernel] + argmax kernel per step"})
    line = {"metric": "decode tokens/sec %s Q4_0 bs=1" % mname, "value": tps, "unit": "tokens/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_value / steps, "higher_is_better": True,
            "scaling": "weak" if (world > 1 and not tp) else "strong", "vs_baseline": None,
            "dtype": "int4 x int4 -> int32 block dots (dp4a), fp32 lane accumulation (bit-exact AVX2 order), f32 KV",
            "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": jobs * steps / e2e_s, "unit": "tokens/s", "h2d_bytes_per_step": 4, "d2h_bytes_per_step": n_vocab * 4,
                    "note": "b200_llama_eval per token: token id by value, logits to pinned host memory, host arg-max"},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "replicas": replicas}
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if saved_stdout is not None:
        os.dup2(2, 1)
    model.free()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
