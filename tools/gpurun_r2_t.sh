#!/bin/bash
# round 2, call T: tc kernel, one proxy fence per quad (8 operand buffers) -- parity + timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_batch.py -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log
tail -3 gpurun_out/r2t_pytest.log
timeout 300 python tools/prompt_probe.py --layers 2 --n 256 --reps 3 2>&1 | tail -1
timeout 600 python -c "import bench; bench.ensure_model(32)" > gpurun_out/model.log 2>&1
timeout 600 python - > gpurun_out/r2t_prompt.log 2>&1 <<'PY'
import time, numpy as np, bench
import llama_swift_b200 as lsb
path = bench.model_path(32)
m = lsb.llama_model_load(path, n_ctx=2100)
rng = np.random.default_rng(0)
for n in (24, 64, 256, 512, 2048):
    toks = rng.integers(3, 32000, size=n).astype(np.int32)
    lsb.llama_eval(m, 8, 0, toks)
    t0 = time.perf_counter(); lsb.llama_eval(m, 8, 0, toks); dt = time.perf_counter() - t0
    print(f"N={n}: {dt*1e3:.2f} ms  {n/dt:.0f} prompt tok/s  (device {m.last_eval_ms:.2f} ms) launches {m.last_launches}", flush=True)
PY
cat gpurun_out/r2t_prompt.log
