#!/bin/bash
# round 2, call J: abort/recover test, query-tiled batch attention (parity + timing)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py "tests/test_gpu_parity.py::test_bounded_wait_abort_and_recover" tests/test_gpu_parity.py::test_llama_eval_vs_oracle -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -5 gpurun_out/r2j_pytest.log
timeout 600 python -c "import bench; bench.ensure_model(32)" > gpurun_out/model.log 2>&1
timeout 600 python - > gpurun_out/r2j_prompt.log 2>&1 <<'PY'
import time, numpy as np, bench
import llama_swift_b200 as lsb
path = bench.model_path(32)
m = lsb.llama_model_load(path, n_ctx=2100)
rng = np.random.default_rng(0)
for tile in (1, 0):
    m.set_option("attn_tile", tile)
    for n in (9, 64, 256, 512, 2048):
        toks = rng.integers(3, 32000, size=n).astype(np.int32)
        lsb.llama_eval(m, 8, 0, toks)
        t0 = time.perf_counter(); lsb.llama_eval(m, 8, 0, toks); dt = time.perf_counter() - t0
        print(f"attn_tile={tile} N={n}: {dt*1e3:.2f} ms  {n/dt:.0f} prompt tok/s  launches {m.last_launches}", flush=True)
PY
cat gpurun_out/r2j_prompt.log
