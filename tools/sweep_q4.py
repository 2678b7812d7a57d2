"""BASELINE.json configs[4]: Q4_1 vs Q4_0 mat-mul kernels over the four LLaMA-7B weight shapes, bs in {1, 4, 16} columns.

Q4_0: bs = 1 is the single-column mat-vec kernel; bs = 4 / 16 are ONE launch of the batch path's mat-mul -- the CUDA-core
multi-column loop (weights streamed once per 8 columns) and the tcgen05 / TMEM kernel (one 16-token tile).  Q4_1 follows
the reference's scalar-only kernel (one sequential float chain per row, ggml.c:1584-1626) and has no multi-column form:
bs > 1 is bs back-to-back launches.  Prints a markdown table (development aid / profiles)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_swift_b200 as lsb
from llama_swift_b200 import ggml_format as gf

rng = np.random.default_rng(0)
print("| shape (M x K) | type | bytes | bs=1 us | GB/s | bs=4 us (cols / tcgen05) | bs=16 us (cols / tcgen05) |")
print("|---|---|---|---|---|---|---|")
for (M, K) in [(4096, 4096), (11008, 4096), (4096, 11008), (32000, 4096)]:
    w = (rng.standard_normal((M, K)) / np.sqrt(K)).astype(np.float32)
    x = rng.standard_normal((16, K)).astype(np.float32)
    blk = gf.quantize_q4_0(w)
    _, ms1 = lsb.q4_0_matvec(blk, x[0], timed=True)
    cells = []
    for n in (4, 16):
        _, a = lsb.q4_0_matmul(blk, x[:n], path=0, timed=True)
        _, b = lsb.q4_0_matmul(blk, x[:n], path=1, timed=True)
        cells.append(f"{a * 1e3:.1f} / {b * 1e3:.1f}")
    nbytes = M * K // 32 * 20
    print(f"| {M} x {K} | Q4_0 | {nbytes / 1e6:.1f} MB | {ms1 * 1e3:.1f} | {nbytes / ms1 / 1e6:.0f} | {cells[0]} | {cells[1]} |", flush=True)
    blk1 = gf.quantize_q4_1(w)
    _, ms = lsb.q4_1_matvec(blk1, x[0], timed=True)
    nbytes = M * K // 32 * 24
    print(f"| {M} x {K} | Q4_1 | {nbytes / 1e6:.1f} MB | {ms * 1e3:.1f} | {nbytes / ms / 1e6:.0f} | {4 * ms * 1e3:.1f} (4 launches) | {16 * ms * 1e3:.1f} (16 launches) |", flush=True)
