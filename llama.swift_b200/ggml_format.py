"""Writer for the reference's on-disk model format ("ggml" magic, un-versioned).

The format is defined by the reference's offline tools, which this module restates so that
synthetic models (there are no real weights and no network here) can be produced bit-compatibly:

  * header / vocab / tensor records ........ tools/convert-pth-to-ggml.py:92-169
  * which tensors get quantized, ftype ..... Sources/cpp/quantize.cpp:60-200  (every 2-D "*weight")
  * Q4_0 / Q4_1 offline quantizers ......... Sources/cpp/utils.cpp:431-544    (scalar, round-half-away)
  * multi-part split (columns vs rows) ..... bridge/LlamaPredictOperation.mm:358-388

The loader that consumes these files is llama_model_load (PO.mm:98-498) in the reference and
b200_llama_load in this repository.
"""
from __future__ import annotations

import os
import struct
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import numpy as np

GGML_MAGIC = 0x67676D6C          # PO.mm:110
QK = 32                          # ggml.c:360
FTYPE_F32, FTYPE_F16, FTYPE_Q4_0, FTYPE_Q4_1 = 0, 1, 2, 3   # PO.mm:169-173
LLAMA_N_PARTS = {4096: 1, 5120: 2, 6656: 4, 8192: 8}        # PO.mm:33-38


@dataclass
class HParams:
    n_vocab: int = 32000
    n_embd: int = 4096
    n_mult: int = 256
    n_head: int = 32
    n_layer: int = 32
    ftype: int = FTYPE_Q4_0

    @property
    def n_rot(self) -> int:      # "rot (obsolete)", convert-pth-to-ggml.py:98
        return self.n_embd // self.n_head

    @property
    def n_ff(self) -> int:       # PO.mm:135
        return ((2 * (4 * self.n_embd) // 3 + self.n_mult - 1) // self.n_mult) * self.n_mult


def _round_half_away_f32(v: np.ndarray) -> np.ndarray:
    """C round() on f32 data without going through f64: trunc, then step away from zero when |frac| >= 0.5.
    (v - trunc(v) is exact in f32, so this has no double-rounding problem, unlike floor(|v| + 0.5) in f32.)"""
    t = np.trunc(v)
    frac = v - t
    t += np.where(np.abs(frac) >= np.float32(0.5), np.sign(v), np.float32(0.0)).astype(np.float32)
    return t


def quantize_q4_0(w: np.ndarray) -> np.ndarray:
    """utils.cpp:431-486.  w: [rows, k] f32 -> uint8 [rows, k/32, 20] (f32 d, 16 nibble bytes)."""
    rows, k = w.shape
    assert k % QK == 0
    x = np.ascontiguousarray(w, dtype=np.float32).reshape(rows, k // QK, QK)
    amax = np.max(np.abs(x), axis=2)
    d = (amax / np.float32(7.0)).astype(np.float32)
    with np.errstate(divide="ignore"):
        idv = np.where(d != 0, np.float32(1.0) / d, np.float32(0.0)).astype(np.float32)
    v = x * idv[:, :, None]
    r = _round_half_away_f32(v)                          # C round(): half away from zero
    q = (r.astype(np.int8) + np.int8(8)).view(np.uint8)
    out = np.empty((rows, k // QK, 20), dtype=np.uint8)
    out[:, :, :4] = d.view(np.uint8).reshape(rows, k // QK, 4)
    np.bitwise_or(q[:, :, 0::2], q[:, :, 1::2] << 4, out=out[:, :, 4:])
    return out


def quantize_q4_1(w: np.ndarray) -> np.ndarray:
    """utils.cpp:488-544.  Per-row SoA [nb f32 min][nb f32 d][nb*16 B] -> uint8 [rows, k/32*24]."""
    rows, k = w.shape
    assert k % QK == 0
    nb = k // QK
    x = np.ascontiguousarray(w, dtype=np.float32).reshape(rows, nb, QK)
    mn = np.minimum(np.min(x, axis=2), np.finfo(np.float32).max).astype(np.float32)
    # quirk kept: max starts at numeric_limits<float>::min() (smallest positive normal), utils.cpp:509
    mx = np.maximum(np.max(x, axis=2), np.finfo(np.float32).tiny).astype(np.float32)
    d = ((mx - mn).astype(np.float32) / np.float32(15.0)).astype(np.float32)
    with np.errstate(all="ignore"):
        idv = np.where(d != 0, np.float32(1.0) / d, np.float32(0.0)).astype(np.float32)
        v = ((x - mn[:, :, None]).astype(np.float32) * idv[:, :, None]).astype(np.float32)
        r = _round_half_away_f32(v)
        q = r.astype(np.int64).astype(np.uint8)
    packed = (q[:, :, 0::2] | (q[:, :, 1::2] << 4)).astype(np.uint8)
    out = np.empty((rows, nb * 24), dtype=np.uint8)
    out[:, : nb * 4] = mn.view(np.uint8).reshape(rows, nb * 4)
    out[:, nb * 4: nb * 8] = d.view(np.uint8).reshape(rows, nb * 4)
    out[:, nb * 8:] = packed.reshape(rows, nb * 16)
    return out


def dequantize_q4_0(blocks: np.ndarray) -> np.ndarray:
    """ggml.c:651-684.  uint8 [rows, nb, 20] -> f32 [rows, nb*32]."""
    rows, nb, _ = blocks.shape
    d = blocks[:, :, :4].copy().view(np.float32).reshape(rows, nb)
    qs = blocks[:, :, 4:]
    lo = (qs & 0xF).astype(np.int32) - 8
    hi = (qs >> 4).astype(np.int32) - 8
    q = np.empty((rows, nb, 32), dtype=np.float32)
    q[:, :, 0::2] = lo
    q[:, :, 1::2] = hi
    return (q * d[:, :, None]).astype(np.float32).reshape(rows, nb * 32)


def _direct_q4_0(rng: np.random.Generator, rows: int, k: int, std: float) -> np.ndarray:
    """Synthesize Q4_0 blocks directly (fast path for 7B/13B-sized benchmark files).  Each nibble is
    1 + (a & 7) + (b & 7) for two random bytes a, b: a triangular distribution over 1..15 centred on 8, i.e.
    weights q-8 in -7..7 with std 3.24 like a real quantized tensor (never the unused value 0).  The per-block
    scale is log-normal around the value that gives the requested weight std.  Same byte layout as quantize_q4_0."""
    nb = k // QK
    out = np.empty((rows, nb, 20), dtype=np.uint8)
    a = rng.integers(0, 256, size=(rows, nb, 16), dtype=np.uint8)
    b = rng.integers(0, 256, size=(rows, nb, 16), dtype=np.uint8)
    out[:, :, 4:] = (a & 0x77) + (b & 0x77) + 0x11
    d = (std / 3.24 * np.exp(0.1 * rng.standard_normal((rows, nb), dtype=np.float32))).astype(np.float32)
    out[:, :, :4] = d.view(np.uint8).reshape(rows, nb, 4)
    return out


def tensor_names(hp: HParams):
    """(name, rows, cols, split) in checkpoint order.  split: 0 = by columns (ne[0]), 1 = by rows (ne[1]),
    None = 1-D replicated (PO.mm:358-388)."""
    e, f, v = hp.n_embd, hp.n_ff, hp.n_vocab
    yield ("tok_embeddings.weight", v, e, 0)
    yield ("norm.weight", e, None, None)
    yield ("output.weight", v, e, 1)
    for i in range(hp.n_layer):
        p = f"layers.{i}."
        yield (p + "attention.wq.weight", e, e, 1)
        yield (p + "attention.wk.weight", e, e, 1)
        yield (p + "attention.wv.weight", e, e, 1)
        yield (p + "attention.wo.weight", e, e, 0)
        yield (p + "feed_forward.w1.weight", f, e, 1)
        yield (p + "feed_forward.w2.weight", e, f, 0)
        yield (p + "feed_forward.w3.weight", f, e, 1)
        yield (p + "attention_norm.weight", e, None, None)
        yield (p + "ffn_norm.weight", e, None, None)


def default_vocab(n_vocab: int):
    """ids 0..2 are empty strings like the control tokens at convert-pth-to-ggml.py:108-110; the rest are
    distinct dummy pieces (the greedy tokenizer at utils.cpp:275-311 only needs distinct strings)."""
    words = [b"", b"", b""]
    for i in range(3, n_vocab):
        words.append((" t%d" % i).encode())
    return words[:n_vocab]


def write_header(f, hp: HParams, vocab) -> None:
    f.write(struct.pack("<i", GGML_MAGIC))
    f.write(struct.pack("<7i", hp.n_vocab, hp.n_embd, hp.n_mult, hp.n_head, hp.n_layer, hp.n_rot, hp.ftype))
    assert len(vocab) == hp.n_vocab
    for w in vocab:
        f.write(struct.pack("<i", len(w)))
        f.write(w)


def write_tensor_header(f, name: str, n_dims: int, ftype: int, ne) -> None:
    sname = name.encode()
    f.write(struct.pack("<iii", n_dims, len(sname), ftype))
    for d in ne:
        f.write(struct.pack("<i", d))
    f.write(sname)


def write_synthetic_model(path: str, hp: HParams, seed: int = 0, mode: str = "quantize",
                          resid_scale: float = 0.02, n_parts: int | None = None, vocab=None) -> dict:
    """Write a synthetic model in the reference's format.

    mode "quantize": Gaussian f32 weights ~ N(0, 1/fan_in) rounded through f16 (the converter stores f16,
        convert-pth-to-ggml.py:154-159) then quantized by the restated offline quantizer -- the faithful path,
        used by the parity tests.
    mode "direct":  Q4_0 blocks synthesized directly (benchmark-sized files; seconds instead of minutes).
    wo / w2 are scaled by resid_scale so the network is residual-dominant like trained weights (random
    full-scale weights make the reference chaotic under its own 4-bit activation re-quantization).
    n_parts defaults to what the loader will look for: LLAMA_N_PARTS[n_embd] (PO.mm:136).
    Returns {"files": [...], "bytes": total}.
    """
    if n_parts is None:
        n_parts = LLAMA_N_PARTS[hp.n_embd]
    assert hp.ftype in (FTYPE_Q4_0, FTYPE_Q4_1)
    if mode == "direct":
        assert hp.ftype == FTYPE_Q4_0
    vocab = vocab if vocab is not None else default_vocab(hp.n_vocab)
    files = [path if p == 0 else f"{path}.{p}" for p in range(n_parts)]
    outs = [open(fn, "wb") for fn in files]
    total = 0
    pool = ThreadPoolExecutor(max_workers=max(1, min(32, os.cpu_count() or 1)))
    try:
        for f in outs:
            write_header(f, hp, vocab)
        for name, rows, cols, split in tensor_names(hp):
            key = [seed] + list(name.encode())
            if cols is None:   # 1-D norm weight, f32, replicated in every part (PO.mm:446-457)
                rng = np.random.default_rng(key)
                w = (1.0 + 0.1 * rng.standard_normal(rows)).astype(np.float32)
                for f in outs:
                    write_tensor_header(f, name, 1, FTYPE_F32, [rows])
                    f.write(w.tobytes())
                continue
            std = 1.0 / np.sqrt(cols)
            if name.endswith("wo.weight") or name.endswith("w2.weight"):
                std *= resid_scale
            if name.startswith("tok_embeddings"):
                std = 1.0
            for p, f in enumerate(outs):
                pr, pc = (rows, cols // n_parts) if split == 0 else (rows // n_parts, cols)
                write_tensor_header(f, name, 2, hp.ftype, [pc, pr])
            # row chunks are generated (and quantized) in parallel; each chunk has its own seeded stream, so the
            # file does not depend on the worker count
            chunk = max(1, (16 << 20) // (cols * 4))
            ranges = [(r0, min(rows, r0 + chunk)) for r0 in range(0, rows, chunk)]

            def make(rr, key=key, std=std, cols=cols):
                r0, r1 = rr
                rng = np.random.default_rng(key + [r0])
                if mode == "direct":
                    return None, _direct_q4_0(rng, r1 - r0, cols, std).reshape(r1 - r0, -1)
                w = (rng.standard_normal((r1 - r0, cols), dtype=np.float32) * np.float32(std))
                w = w.astype(np.float16).astype(np.float32)
                if hp.ftype == FTYPE_Q4_0:
                    return None, quantize_q4_0(w).reshape(r1 - r0, -1)
                return w, None

            for (r0, r1), (w, q) in zip(ranges, pool.map(make, ranges)):
                for p, f in enumerate(outs):
                    if split == 1:
                        # rows [p*rows/n_parts, (p+1)*rows/n_parts) live in part p (PO.mm:478-487)
                        lo, hi = p * (rows // n_parts), (p + 1) * (rows // n_parts)
                        a, b = max(r0, lo), min(r1, hi)
                        if a >= b:
                            continue
                        if hp.ftype == FTYPE_Q4_0:
                            f.write(q[a - r0: b - r0].tobytes())
                        else:
                            f.write(quantize_q4_1(w[a - r0: b - r0]).tobytes())
                    else:
                        # columns [p*cols/n_parts, ...) of every row live in part p (PO.mm:467-477)
                        c0, c1 = p * (cols // n_parts), (p + 1) * (cols // n_parts)
                        if hp.ftype == FTYPE_Q4_0:
                            f.write(np.ascontiguousarray(q[:, c0 // QK * 20: c1 // QK * 20]).tobytes())
                        else:
                            # Q4_1 rows are SoA per row, so a column slice is quantized on its own
                            f.write(quantize_q4_1(w[:, c0:c1]).tobytes())
        for f in outs:
            total += f.tell()
    finally:
        pool.shutdown()
        for f in outs:
            f.close()
    return {"files": files, "bytes": total}
