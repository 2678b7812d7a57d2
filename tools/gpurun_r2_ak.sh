#!/bin/bash
# round 2, call AK: Q4_1 term math, packed variant without the contractible mul -> add pair (-DB200_Q41_F32X2=2): parity + timing
mkdir -p gpurun_out
cat > /tmp/q41_time.py <<'PY'
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import llama_swift_b200 as lsb
from llama_swift_b200 import ggml_format as gf
rng = np.random.default_rng(0)
for M, K in ((4096, 4096), (4096, 11008)):
    w = gf.quantize_q4_1((rng.standard_normal((M, K)) / np.sqrt(K)).astype(np.float32))
    x = rng.standard_normal(K).astype(np.float32)
    out, ms = lsb.q4_1_matvec(w, x, timed=True)
    print(f"{M} x {K}: {ms * 1e3:.1f} us", flush=True)
PY
for lib in libb200_q41x2.so libb200llama.so; do
  echo "== $lib"
  B200_LIB=$PWD/llama.swift_b200/$lib timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "q4_1" 2>&1 | tail -1
  B200_LIB=$PWD/llama.swift_b200/$lib timeout 300 python /tmp/q41_time.py
done
