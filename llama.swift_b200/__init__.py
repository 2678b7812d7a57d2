"""llama.swift_b200 -- B200-native (sm_100a) replacement for llama.swift's ggml Q4_0 decode hot path.

Python host-side mirror of the reference's driver interface (bridge/LlamaPredictOperation.mm): the same two entry
points with the same argument meaning and error behaviour --

    llama_model_load(fname, n_ctx)                      PO.mm:98    -> LlamaModel
    llama_eval(model, n_threads, n_past, embd_inp)      PO.mm:510   -> logits of the last token

-- implemented by calling the C ABI of libb200llama.so (include/b200_llama.h) through ctypes.  There is no CPU
path: if the CUDA library is missing or no GPU is present the calls raise.  (The directory name contains a dot, so
import it through the repo-root shim:  `import llama_swift_b200`.)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import ggml_format  # noqa: F401  (writer for the reference's model file format)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200_LIB") or os.path.join(_HERE, "libb200llama.so")   # B200_LIB: development A/B builds

ERR_LOAD = -1000      # LlamaErrorCodeFailedToLoadModel, headers/LlamaError.h:17
ERR_PREDICT = -1001   # LlamaErrorCodePredictionFailed,  headers/LlamaError.h:18


class LlamaError(RuntimeError):
    """Mirror of NSError(domain com.alexrozanski.llama.error, code) raised by the bridge (PO.mm:90-95)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code
        self.message = message


class RunParams(C.Structure):
    """b200_run_params: gpt_params (utils.h:15-37) as _LlamaRunnerBridge fills it (LlamaRunnerBridge.mm:34-43)."""
    _fields_ = [("seed", C.c_int), ("n_threads", C.c_int), ("n_predict", C.c_int), ("repeat_last_n", C.c_int), ("top_k", C.c_int),
                ("top_p", C.c_float), ("temp", C.c_float), ("repeat_penalty", C.c_float), ("n_batch", C.c_int), ("n_ctx", C.c_int),
                ("device", C.c_int)]


EVENT_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_char), C.c_int, C.c_int)
EVAL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_float),
                      C.POINTER(C.c_char), C.c_size_t)
(EVENT_STARTED_LOADING_MODEL, EVENT_FINISHED_LOADING_MODEL, EVENT_STARTED_GENERATING_OUTPUT, EVENT_OUTPUT_TOKEN, EVENT_COMPLETED,
 EVENT_FAILED) = range(6)

_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library.  Fails loudly when it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python llama.swift_b200/build.py` "
                          "(nvcc, sm_100a).  llama.swift_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, ci, cp, sz = C.c_void_p, C.c_int, C.c_char_p, C.c_size_t
    L.b200_llama_load.argtypes = [cp, ci, ci, C.POINTER(vp), cp, sz]
    L.b200_llama_load.restype = ci
    if "B200_LIB" not in os.environ or hasattr(L, "b200_llama_load_group"):   # (older development A/B builds lack the group API)
        L.b200_llama_load_group.argtypes = [cp, ci, C.POINTER(ci), ci, C.POINTER(vp), cp, sz]
        L.b200_llama_load_group.restype = ci
        L.b200_llama_load_shard.argtypes = [cp, ci, ci, ci, ci, C.POINTER(vp), cp, sz]
        L.b200_llama_load_shard.restype = ci
        L.b200_llama_tp_ipc_handle.argtypes = [vp, vp, sz]
        L.b200_llama_tp_ipc_handle.restype = ci
        L.b200_llama_tp_connect_ipc.argtypes = [vp, vp, sz, cp, sz]
        L.b200_llama_tp_connect_ipc.restype = ci
    L.b200_llama_eval.argtypes = [vp, ci, ci, vp, ci, vp, cp, sz]
    L.b200_llama_eval.restype = ci
    L.b200_llama_free.argtypes = [vp]
    L.b200_llama_free.restype = None
    for n in ("n_vocab", "n_ctx", "n_embd", "n_layer", "n_head", "ftype"):
        f = getattr(L, "b200_llama_" + n)
        f.argtypes, f.restype = [vp], ci
    L.b200_llama_token_str.argtypes = [vp, ci, C.POINTER(ci)]
    L.b200_llama_token_str.restype = C.POINTER(C.c_char)
    L.b200_llama_decode_device.argtypes = [vp, ci, ci, ci, ci, vp, vp, vp, C.POINTER(C.c_float), cp, sz]
    L.b200_llama_decode_device.restype = ci
    L.b200_llama_kv_export.argtypes = [vp, ci, ci, ci, vp]
    L.b200_llama_kv_export.restype = ci
    L.b200_llama_kv_import.argtypes = [vp, ci, ci, ci, vp]
    L.b200_llama_kv_import.restype = ci
    L.b200_llama_last_launches.argtypes, L.b200_llama_last_launches.restype = [vp], C.c_longlong
    L.b200_llama_weight_bytes.argtypes, L.b200_llama_weight_bytes.restype = [vp], C.c_longlong
    L.b200_llama_last_kernel_ms.argtypes, L.b200_llama_last_kernel_ms.restype = [vp], C.c_double
    if "B200_LIB" not in os.environ or hasattr(L, "b200_llama_last_eval_ms"):
        L.b200_llama_last_eval_ms.argtypes, L.b200_llama_last_eval_ms.restype = [vp], C.c_double
    L.b200_llama_set_option.argtypes, L.b200_llama_set_option.restype = [vp, cp, ci], ci
    L.b200_q4_0_matvec.argtypes = [ci, vp, ci, ci, vp, vp, ci, C.POINTER(C.c_float), cp, sz]
    L.b200_q4_0_matvec.restype = ci
    if "B200_LIB" not in os.environ or hasattr(L, "b200_q4_0_matmul"):
        L.b200_q4_0_matmul.argtypes = [ci, vp, ci, ci, vp, ci, vp, ci, C.POINTER(C.c_float), cp, sz]
        L.b200_q4_0_matmul.restype = ci
    L.b200_q4_1_matvec.argtypes = [ci, vp, ci, ci, vp, vp, C.POINTER(C.c_float), cp, sz]
    L.b200_q4_1_matvec.restype = ci
    if "B200_LIB" not in os.environ or hasattr(L, "b200_llama_acquire"):
        L.b200_llama_acquire.argtypes, L.b200_llama_acquire.restype = [cp, ci, ci, C.POINTER(vp), cp, sz], ci
        L.b200_llama_release.argtypes, L.b200_llama_release.restype = [vp], None
        L.b200_llama_cache_clear.argtypes, L.b200_llama_cache_clear.restype = [], None
    if "B200_LIB" not in os.environ or hasattr(L, "b200_llama_tokenize"):
        L.b200_tokenizer_create.argtypes, L.b200_tokenizer_create.restype = [vp], vp
        L.b200_tokenizer_create_from.argtypes, L.b200_tokenizer_create_from.restype = [C.POINTER(cp), C.POINTER(ci), ci], vp
        L.b200_tokenizer_free.argtypes, L.b200_tokenizer_free.restype = [vp], None
        L.b200_llama_tokenize.argtypes, L.b200_llama_tokenize.restype = [vp, cp, sz, ci, vp, ci], ci
        L.b200_rng_create.argtypes, L.b200_rng_create.restype = [ci], vp
        L.b200_rng_free.argtypes, L.b200_rng_free.restype = [vp], None
        L.b200_llama_sample_top_p_top_k.argtypes = [ci, vp, vp, ci, C.c_double, ci, C.c_double, C.c_double, vp]
        L.b200_llama_sample_top_p_top_k.restype = ci
    if "B200_LIB" not in os.environ or hasattr(L, "b200_llama_eval_topk"):
        dbl = C.c_double
        L.b200_llama_eval_topk.argtypes = [vp, ci, ci, vp, ci, vp, ci, dbl, dbl, ci, vp, vp, C.POINTER(ci), cp, sz]
        L.b200_llama_eval_topk.restype = ci
        L.b200_llama_last_logits.argtypes, L.b200_llama_last_logits.restype = [vp, vp, cp, sz], ci
        L.b200_llama_sample_from_candidates.argtypes = [vp, vp, ci, dbl, vp]
        L.b200_llama_sample_from_candidates.restype = ci
        L.b200_sample_topk.argtypes = [ci, vp, ci, vp, ci, dbl, dbl, ci, vp, vp, C.POINTER(ci), C.POINTER(C.c_float), cp, sz]
        L.b200_sample_topk.restype = ci
        L.b200_llama_run_sampler_stats.argtypes = [C.POINTER(ci), C.POINTER(ci)]
        L.b200_llama_run_sampler_stats.restype = None
    if "B200_LIB" not in os.environ or hasattr(L, "b200_llama_run"):
        L.b200_run_params_default.argtypes, L.b200_run_params_default.restype = [C.POINTER(RunParams)], None
        L.b200_llama_run.argtypes = [cp, cp, sz, cp, sz, C.POINTER(RunParams), EVENT_FN, vp]
        L.b200_llama_run.restype = ci
        L.b200_llama_run_loop.argtypes = [EVAL_FN, vp, ci, ci, vp, C.POINTER(cp), C.POINTER(ci), cp, sz, cp, sz,
                                          C.POINTER(RunParams), EVENT_FN, vp]
        L.b200_llama_run_loop.restype = ci
    _lib = L
    return L


class LlamaModel:
    """llama_model + gpt_vocab as the token loop sees them (PO.mm:71-88, utils.h:49-55)."""

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)
        L = lib()
        self.n_vocab = L.b200_llama_n_vocab(self._h)
        self.n_ctx = L.b200_llama_n_ctx(self._h)
        self.n_embd = L.b200_llama_n_embd(self._h)
        self.n_layer = L.b200_llama_n_layer(self._h)
        self.n_head = L.b200_llama_n_head(self._h)
        self.ftype = L.b200_llama_ftype(self._h)

    _cached = False

    def free(self) -> None:            # ggml_free(model.ctx), PO.mm:900
        if self._h:
            if self._cached:
                lib().b200_llama_release(self._h)      # stays resident for the next run
            else:
                lib().b200_llama_free(self._h)
            self._h = None

    release = free

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def id_to_token(self, i: int) -> bytes:
        n = C.c_int(0)
        p = lib().b200_llama_token_str(self._h, i, C.byref(n))
        return C.string_at(p, n.value)

    def set_option(self, key: str, value: int) -> None:
        if lib().b200_llama_set_option(self._h, key.encode(), value) != 0:
            raise KeyError(key)

    @property
    def last_launches(self) -> int:
        return int(lib().b200_llama_last_launches(self._h))

    @property
    def kernel_ms_total(self) -> float:
        """Sum of the per-launch token-kernel durations of the last decode_device call (option time_kernel = 1)."""
        return float(lib().b200_llama_last_kernel_ms(self._h))

    @property
    def last_eval_ms(self) -> float:
        """Device time (CUDA events) of the last batched llama_eval (n_tokens > 1)."""
        return float(lib().b200_llama_last_eval_ms(self._h))

    @property
    def weight_bytes(self) -> int:
        return int(lib().b200_llama_weight_bytes(self._h))

    def kv_export(self, layer: int, which: int, n_rows: int) -> np.ndarray:
        out = np.empty((n_rows, self.n_embd), dtype=np.float32)
        if lib().b200_llama_kv_export(self._h, layer, which, n_rows, out.ctypes.data) != 0:
            raise LlamaError(ERR_PREDICT, "kv_export failed")
        return out

    def kv_import(self, layer: int, which: int, rows: np.ndarray) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if lib().b200_llama_kv_import(self._h, layer, which, rows.shape[0], rows.ctypes.data) != 0:
            raise LlamaError(ERR_PREDICT, "kv_import failed")

    def decode_device(self, n_past: int, first_token: int, n_steps: int, n_threads: int = 8, forced_tokens=None,
                      want_logits: bool = False):
        """Device-resident greedy / teacher-forced loop.  Returns (argmax tokens, logits or None, elapsed ms)."""
        toks = np.empty(n_steps, dtype=np.int32)
        forced = None if forced_tokens is None else np.ascontiguousarray(forced_tokens, dtype=np.int32)
        logits = np.empty((n_steps, self.n_vocab), dtype=np.float32) if want_logits else None
        ms = C.c_float(0)
        err = C.create_string_buffer(512)
        rc = lib().b200_llama_decode_device(self._h, n_threads, n_past, int(first_token), n_steps,
                                            None if forced is None else forced.ctypes.data, toks.ctypes.data,
                                            None if logits is None else logits.ctypes.data, C.byref(ms), err, 512)
        if rc != 0:
            raise LlamaError(rc, err.value.decode(errors="replace"))
        return toks, logits, ms.value


def llama_model_load(fname: str, n_ctx: int = 512, device: int = 0) -> LlamaModel:
    """llama_model_load(fname, model, vocab, n_ctx, &err), PO.mm:98.  The reference's caller passes 512 (PO.mm:790)."""
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = lib().b200_llama_load(os.fsencode(fname), n_ctx, device, C.byref(h), err, 512)
    if rc != 0:
        raise LlamaError(rc, err.value.decode(errors="replace"))
    return LlamaModel(h.value)


def llama_model_acquire(fname: str, n_ctx: int = 512, device: int = 0) -> LlamaModel:
    """Like llama_model_load, but a model that a previous run released is handed back without touching the file
    (the reference re-loads on every run(), PO.mm:790).  Give it back with .release()."""
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = lib().b200_llama_acquire(os.fsencode(fname), n_ctx, device, C.byref(h), err, 512)
    if rc != 0:
        raise LlamaError(rc, err.value.decode(errors="replace"))
    m = LlamaModel(h.value)
    m._cached = True
    return m


def llama_model_cache_clear() -> None:
    lib().b200_llama_cache_clear()


def llama_model_load_group(fname: str, n_ctx: int = 512, devices=(0, 1)) -> LlamaModel:
    """One model over several GPUs driven by this process (tensor parallel, rows of every matrix split over the
    group; include/b200_llama.h).  The returned handle is used exactly like a single-GPU one."""
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    devs = (C.c_int * len(devices))(*devices)
    rc = lib().b200_llama_load_group(os.fsencode(fname), n_ctx, devs, len(devices), C.byref(h), err, 512)
    if rc != 0:
        raise LlamaError(rc, err.value.decode(errors="replace"))
    return LlamaModel(h.value)


IPC_HANDLE_BYTES = 64


def llama_model_load_shard(fname: str, n_ctx: int, device: int, tp_rank: int, tp_size: int) -> LlamaModel:
    """Rank tp_rank of a one-process-per-GPU tensor-parallel group; connect it with tp_connect() before evaluating."""
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = lib().b200_llama_load_shard(os.fsencode(fname), n_ctx, device, tp_rank, tp_size, C.byref(h), err, 512)
    if rc != 0:
        raise LlamaError(rc, err.value.decode(errors="replace"))
    return LlamaModel(h.value)


def tp_ipc_handle(model: LlamaModel) -> bytes:
    buf = C.create_string_buffer(IPC_HANDLE_BYTES)
    if lib().b200_llama_tp_ipc_handle(model._h, buf, IPC_HANDLE_BYTES) != 0:
        raise LlamaError(ERR_LOAD, "cudaIpcGetMemHandle failed")
    return buf.raw


def tp_connect(model: LlamaModel, handles) -> None:
    """handles: every rank's tp_ipc_handle(), in rank order (exchange them with torch.distributed.all_gather_object
    or any other channel -- that exchange is plumbing and happens once, at load time)."""
    blob = b"".join(bytes(h) for h in handles)
    buf = C.create_string_buffer(blob, len(blob))
    err = C.create_string_buffer(512)
    rc = lib().b200_llama_tp_connect_ipc(model._h, buf, IPC_HANDLE_BYTES, err, 512)
    if rc != 0:
        raise LlamaError(rc, err.value.decode(errors="replace"))


def llama_eval(model: LlamaModel, n_threads: int, n_past: int, embd_inp) -> np.ndarray:
    """llama_eval(model, n_threads, n_past, embd_inp, embd_w, mem_per_token, &err), PO.mm:510-518:
    returns embd_w, the n_vocab logits of the LAST token of embd_inp (PO.mm:724-725)."""
    toks = np.ascontiguousarray(embd_inp, dtype=np.int32)
    logits = np.empty(model.n_vocab, dtype=np.float32)
    err = C.create_string_buffer(512)
    rc = lib().b200_llama_eval(model._h, n_threads, n_past, toks.ctypes.data, len(toks), logits.ctypes.data, err, 512)
    if rc != 0:
        raise LlamaError(rc, err.value.decode(errors="replace"))
    return logits


def llama_eval_topk(model: LlamaModel, n_threads: int, n_past: int, embd_inp, last_n_tokens, repeat_penalty=1.3, top_k=40, temp=0.8):
    """llama_eval + the candidate stage of llama_sample_top_p_top_k (utils.cpp:359-386) on the GPU: returns (values, ids), the
    reference's logits_id after sample_top_k -- or None when that order is not determined by the values alone (then
    llama_last_logits + Sampler.sample give the reference's answer)."""
    toks = np.ascontiguousarray(embd_inp, dtype=np.int32)
    last = np.ascontiguousarray(last_n_tokens, dtype=np.int32)
    vals, ids, n = np.empty(max(1, top_k), np.float64), np.empty(max(1, top_k), np.int32), C.c_int(0)
    err = C.create_string_buffer(512)
    rc = lib().b200_llama_eval_topk(model._h, n_threads, n_past, toks.ctypes.data, len(toks), last.ctypes.data, len(last), repeat_penalty,
                                    temp, top_k, vals.ctypes.data, ids.ctypes.data, C.byref(n), err, 512)
    if rc != 0:
        raise LlamaError(rc, err.value.decode(errors="replace"))
    return (vals[:n.value], ids[:n.value]) if n.value else None


def llama_last_logits(model: LlamaModel) -> np.ndarray:
    """The logits of the last evaluation, still on the device after llama_eval_topk."""
    logits = np.empty(model.n_vocab, dtype=np.float32)
    err = C.create_string_buffer(512)
    rc = lib().b200_llama_last_logits(model._h, logits.ctypes.data, err, 512)
    if rc != 0:
        raise LlamaError(rc, err.value.decode(errors="replace"))
    return logits


def sample_topk(logits, last_n_tokens, repeat_penalty=1.3, top_k=40, temp=0.8, device: int = 0, timed: bool = False):
    """Kernel-level entry of the sampler's candidate stage on host logits: (values, ids) or None if ambiguous."""
    lg = np.ascontiguousarray(logits, dtype=np.float32)
    last = np.ascontiguousarray(last_n_tokens, dtype=np.int32)
    vals, ids, n, ms = np.empty(max(1, top_k), np.float64), np.empty(max(1, top_k), np.int32), C.c_int(0), C.c_float(0)
    err = C.create_string_buffer(512)
    rc = lib().b200_sample_topk(device, lg.ctypes.data, len(lg), last.ctypes.data, len(last), repeat_penalty, temp, top_k,
                                vals.ctypes.data, ids.ctypes.data, C.byref(n), C.byref(ms) if timed else None, err, 512)
    if rc != 0:
        raise LlamaError(rc, err.value.decode(errors="replace"))
    res = (vals[:n.value], ids[:n.value]) if n.value else None
    return (res, ms.value) if timed else res


def run_sampler_stats():
    """(sampling steps served with the candidate stage on the GPU, by the host sampler) of this thread's last LlamaRunner.run."""
    a, b = C.c_int(0), C.c_int(0)
    lib().b200_llama_run_sampler_stats(C.byref(a), C.byref(b))
    return a.value, b.value


def q4_0_matvec(w_blocks: np.ndarray, x: np.ndarray, lane_pairs: int = 0, device: int = 0, timed: bool = False):
    """out[M] = W (Q4_0, ggml rows of 20-byte blocks) * x through the production mat-vec kernel."""
    w = np.ascontiguousarray(w_blocks, dtype=np.uint8)
    x = np.ascontiguousarray(x, dtype=np.float32)
    K = x.shape[0]
    M = w.size // (K // 32 * 20)
    out = np.empty(M, dtype=np.float32)
    ms = C.c_float(0)
    err = C.create_string_buffer(512)
    rc = lib().b200_q4_0_matvec(device, w.ctypes.data, M, K, x.ctypes.data, out.ctypes.data, lane_pairs,
                                C.byref(ms) if timed else None, err, 512)
    if rc != 0:
        raise LlamaError(rc, err.value.decode(errors="replace"))
    return (out, ms.value) if timed else out


def q4_0_matmul(w_blocks: np.ndarray, x: np.ndarray, path: int = 0, device: int = 0, timed: bool = False):
    """out[N, M] = W x the N columns x[N, K] (kernel-level entry of the batch path; path 0 CUDA cores, 1 tcgen05)."""
    w = np.ascontiguousarray(w_blocks, dtype=np.uint8)
    x = np.ascontiguousarray(x, dtype=np.float32)
    N, K = x.shape
    M = w.size // (K // 32 * 20)
    out = np.empty((N, M), dtype=np.float32)
    ms = C.c_float(0)
    err = C.create_string_buffer(512)
    rc = lib().b200_q4_0_matmul(device, w.ctypes.data, M, K, x.ctypes.data, N, out.ctypes.data, path,
                                C.byref(ms) if timed else None, err, 512)
    if rc != 0:
        raise LlamaError(rc, err.value.decode(errors="replace"))
    return (out, ms.value) if timed else out


def q4_1_matvec(w_rows: np.ndarray, x: np.ndarray, device: int = 0, timed: bool = False):
    """out[M] = W (Q4_1, ggml per-row layout) * x through the Q4_1 mat-vec kernel."""
    w = np.ascontiguousarray(w_rows, dtype=np.uint8)
    x = np.ascontiguousarray(x, dtype=np.float32)
    K = x.shape[0]
    M = w.size // (K // 32 * 24)
    out = np.empty(M, dtype=np.float32)
    ms = C.c_float(0)
    err = C.create_string_buffer(512)
    rc = lib().b200_q4_1_matvec(device, w.ctypes.data, M, K, x.ctypes.data, out.ctypes.data, C.byref(ms) if timed else None, err, 512)
    if rc != 0:
        raise LlamaError(rc, err.value.decode(errors="replace"))
    return (out, ms.value) if timed else out


class Tokenizer:
    """llama_tokenize(vocab, text, bos), utils.cpp:275-311, over a byte trie (include/b200_llama.h)."""

    def __init__(self, model: LlamaModel = None, pieces=None):
        L = lib()
        if model is not None:
            self._h = C.c_void_p(L.b200_tokenizer_create(model._h))
        else:
            pieces = [bytes(p) for p in pieces]
            arr = (C.c_char_p * len(pieces))(*pieces)
            lens = (C.c_int * len(pieces))(*[len(p) for p in pieces])
            self._h = C.c_void_p(L.b200_tokenizer_create_from(arr, lens, len(pieces)))
        if not self._h:
            raise LlamaError(ERR_LOAD, "tokenizer construction failed")

    def __call__(self, text, bos: bool = True):
        data = text.encode() if isinstance(text, str) else bytes(text)
        cap = len(data) + 2
        out = np.empty(cap, dtype=np.int32)
        n = lib().b200_llama_tokenize(self._h, data, len(data), int(bos), out.ctypes.data, cap)
        if n < 0:
            raise LlamaError(ERR_PREDICT, "tokenize failed")
        return out[:n].copy()

    def __del__(self):
        try:
            if self._h:
                lib().b200_tokenizer_free(self._h)
                self._h = None
        except Exception:
            pass


class Sampler:
    """llama_sample_top_p_top_k (utils.cpp:345-428) with the std::mt19937 of PO.mm:773; defaults are gpt_params' (utils.h:15-37)."""

    def __init__(self, seed: int = -1):
        self._h = C.c_void_p(lib().b200_rng_create(seed))

    def sample(self, logits, last_n_tokens, repeat_penalty=1.3, top_k=40, top_p=0.95, temp=0.8) -> int:
        lg = np.ascontiguousarray(logits, dtype=np.float32)
        last = np.ascontiguousarray(last_n_tokens, dtype=np.int32)
        return int(lib().b200_llama_sample_top_p_top_k(len(lg), lg.ctypes.data, last.ctypes.data, len(last),
                                                       repeat_penalty, top_k, top_p, temp, self._h))

    def sample_from_candidates(self, values, ids, top_p=0.95) -> int:
        """The rest of llama_sample_top_p_top_k (utils.cpp:388-428) on the candidates of llama_eval_topk / sample_topk."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        i = np.ascontiguousarray(ids, dtype=np.int32)
        return int(lib().b200_llama_sample_from_candidates(v.ctypes.data, i.ctypes.data, len(v), top_p, self._h))

    def __del__(self):
        try:
            if self._h:
                lib().b200_rng_free(self._h)
                self._h = None
        except Exception:
            pass


def default_run_params(**overrides) -> RunParams:
    p = RunParams()
    lib().b200_run_params_default(C.byref(p))
    for k, v in overrides.items():
        setattr(p, k, v)
    return p


class LlamaRunner:
    """Mirror of the Swift LlamaRunner (Sources/llama/LlamaRunner.swift:11-124) over b200_llama_run: run(prompt) drives
    the reference's whole token loop (PO.mm:768-901) in the library and hands every event to `on_event(kind, text, code)`;
    returns the list of (id, piece) it emitted (prompt echo included, like the reference)."""

    def __init__(self, model_path: str):
        self.model_path = model_path

    def run(self, prompt: str, params: RunParams = None, reverse_prompt: str = "", on_event=None):
        params = params or default_run_params()
        out, failure = [], []

        def cb(_user, kind, text, n, code):
            piece = C.string_at(text, n) if text else b""
            if kind == EVENT_OUTPUT_TOKEN:
                out.append((code, piece))
            elif kind == EVENT_FAILED:
                failure.append((code, piece.decode(errors="replace")))
            if on_event:
                on_event(kind, piece, code)

        keep = EVENT_FN(cb)
        pb, ab = prompt.encode(), reverse_prompt.encode()
        rc = lib().b200_llama_run(os.fsencode(self.model_path), pb, len(pb), ab, len(ab), C.byref(params), keep, None)
        if rc != 0:
            raise LlamaError(*(failure[0] if failure else (rc, "run failed")))
        return out
