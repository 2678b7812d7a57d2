"""Shared fixtures.  Everything under oracle/ is test infrastructure: it is loaded here (and only by tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs) as the checker for the CUDA path."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import llama_swift_b200 as lsb  # noqa: E402
from llama_swift_b200 import ggml_format as gf  # noqa: E402

ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libllama_ref.so")
CACHE = os.environ.get("B200_TEST_CACHE", "/tmp/b200_llama_test_models")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _bind_model_api(L, prefix):
    vp, ci, cp, sz = C.c_void_p, C.c_int, C.c_char_p, C.c_size_t
    getattr(L, prefix + "_load").restype = vp
    getattr(L, prefix + "_load").argtypes = [cp, ci, cp, sz]
    getattr(L, prefix + "_eval").argtypes = [vp, ci, ci, vp, ci, vp, cp, sz]
    getattr(L, prefix + "_eval").restype = ci
    getattr(L, prefix + "_free").argtypes = [vp]
    getattr(L, prefix + "_free").restype = None
    for n in ("n_vocab", "n_ctx", "n_embd", "n_layer", "n_head"):
        f = getattr(L, f"{prefix}_{n}")
        f.argtypes, f.restype = [vp], ci
    getattr(L, prefix + "_kv_export").argtypes = [vp, ci, ci, ci, vp]
    getattr(L, prefix + "_kv_import").argtypes = [vp, ci, ci, ci, vp]


class CpuModel:
    """Uniform wrapper over the restatement (oracle/liboracle.so) and the compiled reference (oracle/_ref)."""

    def __init__(self, L, prefix, path, n_ctx):
        self.L, self.p = L, prefix
        err = C.create_string_buffer(512)
        self.h = getattr(L, prefix + "_load")(os.fsencode(path), n_ctx, err, 512)
        if not self.h:
            raise RuntimeError(err.value.decode())
        self.h = C.c_void_p(self.h)
        self.n_vocab = getattr(L, prefix + "_n_vocab")(self.h)
        self.n_embd = getattr(L, prefix + "_n_embd")(self.h)
        self.n_layer = getattr(L, prefix + "_n_layer")(self.h)

    def eval(self, n_threads, n_past, tokens):
        toks = np.ascontiguousarray(tokens, dtype=np.int32)
        out = np.empty(self.n_vocab, dtype=np.float32)
        err = C.create_string_buffer(512)
        rc = getattr(self.L, self.p + "_eval")(self.h, n_threads, n_past, toks.ctypes.data, len(toks), out.ctypes.data, err, 512)
        if rc != 0:
            raise RuntimeError(err.value.decode())
        return out

    def kv(self, layer, which, n_rows):
        out = np.empty((n_rows, self.n_embd), dtype=np.float32)
        getattr(self.L, self.p + "_kv_export")(self.h, layer, which, n_rows, out.ctypes.data)
        return out

    def kv_import(self, layer, which, rows):
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        getattr(self.L, self.p + "_kv_import")(self.h, layer, which, rows.shape[0], rows.ctypes.data)

    def free(self):
        if self.h:
            getattr(self.L, self.p + "_free")(self.h)
            self.h = None


@pytest.fixture(scope="session")
def oracle_lib():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ROOT, "oracle", "llama_oracle.c")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    L = C.CDLL(ORACLE_SO)
    _bind_model_api(L, "ora")
    vp, ci = C.c_void_p, C.c_int
    L.ora_quantize_row_q4_0.argtypes = [vp, vp, ci]
    L.ora_quantize_row_q4_1.argtypes = [vp, vp, ci]
    L.ora_mul_mat_q4.argtypes = [ci, vp, ci, ci, vp, ci, vp]
    L.ora_norm.argtypes = [vp, vp, ci]
    L.ora_rope.argtypes = [vp, ci, ci, ci]
    L.ora_soft_max.argtypes = [vp, ci]
    L.ora_silu.argtypes = [vp, vp, ci]
    L.ora_vec_dot_f32.argtypes = [ci, vp, vp]
    L.ora_vec_dot_f32.restype = C.c_float
    return L


@pytest.fixture(scope="session")
def ref_lib():
    """The unmodified reference, compiled by oracle/Makefile (prebuilt .so on the GPU box)."""
    if os.path.isdir("/root/reference/Sources/cpp"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libllama_ref.so not available")
    L = C.CDLL(REF_SO)
    _bind_model_api(L, "ref_llama")
    vp, ci = C.c_void_p, C.c_int
    L.ref_mul_mat_q4.argtypes = [ci, vp, ci, ci, vp, ci, vp, ci]
    L.ref_unary_op.argtypes = [ci, vp, ci, ci, ci, ci, vp, ci]
    L.ref_quantize_row.argtypes = [ci, vp, vp, ci]
    L.ref_dequantize_row.argtypes = [ci, vp, vp, ci]
    L.ref_quantize_weights.argtypes = [ci, vp, vp, ci, ci]
    L.ref_quantize_weights.restype = C.c_size_t
    return L


def model_file(n_layer, n_vocab, seed=1, n_embd=4096, mode="quantize", ftype=2):
    """Synthetic model in the reference's format, cached across tests (Gaussian weights,
    residual-dominant scaling, quantized by the restated offline quantizer)."""
    os.makedirs(CACHE, exist_ok=True)
    n_head = n_embd // 128
    path = os.path.join(CACHE, f"ggml-model-e{n_embd}-l{n_layer}-v{n_vocab}-s{seed}-{mode}-t{ftype}.bin")
    n_parts = gf.LLAMA_N_PARTS[n_embd]
    if not all(os.path.exists(path if p == 0 else f"{path}.{p}") for p in range(n_parts)):
        hp = gf.HParams(n_vocab=n_vocab, n_embd=n_embd, n_head=n_head, n_layer=n_layer, ftype=ftype)
        gf.write_synthetic_model(path + ".tmp", hp, seed=seed, mode=mode)
        for p in range(n_parts):
            os.replace(path + ".tmp" + ("" if p == 0 else f".{p}"), path if p == 0 else f"{path}.{p}")
    return path


@pytest.fixture(scope="session")
def small_model():
    """2 layers of 7B width, 512-token vocabulary (~256 MB)."""
    return model_file(n_layer=2, n_vocab=512)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / max(1e-30, np.linalg.norm(b.astype(np.float64))))
