#!/bin/bash
# round 2, call G (tc kernel v2): tcgen05 prefill mat-mul -- parity + prompt timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_batch.py -m gpu -x -q -s > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
grep -E "\[batch\]|passed|failed|rc=|Error|error" gpurun_out/r2g_pytest.log | tail -40
timeout 600 python -c "import bench; bench.ensure_model(32)" > gpurun_out/model.log 2>&1
timeout 600 python - > gpurun_out/r2g_prompt.log 2>&1 <<'PY'
import time, numpy as np, bench
import llama_swift_b200 as lsb
path = bench.model_path(32)
m = lsb.llama_model_load(path, n_ctx=2100)
rng = np.random.default_rng(0)
for tc in (1, 0):
    m.set_option("tc", tc)
    for n in (4, 9, 64, 256, 512, 2048):
        if tc == 0 and n > 512: continue
        toks = rng.integers(3, 32000, size=n).astype(np.int32)
        lsb.llama_eval(m, 8, 0, toks)
        t0 = time.perf_counter(); lsb.llama_eval(m, 8, 0, toks); dt = time.perf_counter() - t0
        print(f"tc={tc} N={n}: {dt*1e3:.2f} ms  {n/dt:.0f} prompt tok/s  launches {m.last_launches}", flush=True)
PY
cat gpurun_out/r2g_prompt.log
