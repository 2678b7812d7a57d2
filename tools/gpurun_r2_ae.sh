#!/bin/bash
# round 2, call AE: Q4_1 term/chain kernel -- parity tests, sweep, decode speed of a Q4_1 7B-shaped model
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "q4_1" > gpurun_out/r2ae_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ae_pytest.log
tail -4 gpurun_out/r2ae_pytest.log
timeout 600 python tools/sweep_q4.py > gpurun_out/r2ae_sweep_q4.md 2>&1; tail -14 gpurun_out/r2ae_sweep_q4.md
timeout 900 python tools/q4_1_probe.py > gpurun_out/r2ae_q41_probe.log 2>&1; tail -4 gpurun_out/r2ae_q41_probe.log
