#!/bin/bash
# round 2, call I: tc kernel v3 timing, full GPU suite (abort/recover test), compute-sanitizer memcheck + racecheck
mkdir -p gpurun_out
timeout 600 python -c "import bench; bench.ensure_model(32)" > gpurun_out/model.log 2>&1
timeout 600 python - > gpurun_out/r2i_prompt.log 2>&1 <<'PY'
import time, numpy as np, bench
import llama_swift_b200 as lsb
path = bench.model_path(32)
m = lsb.llama_model_load(path, n_ctx=2100)
rng = np.random.default_rng(0)
for n in (4, 9, 64, 256, 512, 2048):
    toks = rng.integers(3, 32000, size=n).astype(np.int32)
    lsb.llama_eval(m, 8, 0, toks)
    t0 = time.perf_counter(); lsb.llama_eval(m, 8, 0, toks); dt = time.perf_counter() - t0
    print(f"tc=1 N={n}: {dt*1e3:.2f} ms  {n/dt:.0f} prompt tok/s  launches {m.last_launches}", flush=True)
PY
cat gpurun_out/r2i_prompt.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -4 gpurun_out/r2i_pytest.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_probe.py > gpurun_out/r2i_memcheck.log 2>&1; tail -6 gpurun_out/r2i_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_probe.py > gpurun_out/r2i_racecheck.log 2>&1; tail -6 gpurun_out/r2i_racecheck.log
