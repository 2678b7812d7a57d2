"""BASELINE.json configs[4]: Q4_1 vs Q4_0 mat-vec kernels over the four LLaMA-7B weight shapes, N in {1, 4, 16} columns.

The reference evaluates the N columns of a mat-mul independently (ggml.c:6199-6222), and so does this library (N
single-column passes), so N > 1 is N back-to-back kernel invocations on the same matrix; what the sweep shows is the
kernel-level cost per column and the achieved weight bandwidth.  Prints a markdown table (development aid / profiles)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_swift_b200 as lsb
from llama_swift_b200 import ggml_format as gf

rng = np.random.default_rng(0)
print("| shape (M x K) | type | bytes | us / column | GB/s | N=4 us | N=16 us |")
print("|---|---|---|---|---|---|---|")
for (M, K) in [(4096, 4096), (11008, 4096), (4096, 11008), (32000, 4096)]:
    w = (rng.standard_normal((M, K)) / np.sqrt(K)).astype(np.float32)
    x = rng.standard_normal(K).astype(np.float32)
    for name, q, fn, bpb in (("Q4_0", gf.quantize_q4_0, lsb.q4_0_matvec, 20), ("Q4_1", gf.quantize_q4_1, lsb.q4_1_matvec, 24)):
        blk = q(w)
        out, ms = fn(blk, x, timed=True)           # best of 5 launches, CUDA events
        nbytes = M * K // 32 * bpb
        print(f"| {M} x {K} | {name} | {nbytes / 1e6:.1f} MB | {ms * 1e3:.1f} | {nbytes / ms / 1e6:.0f} | {4 * ms * 1e3:.1f} | {16 * ms * 1e3:.1f} |", flush=True)
