#!/bin/bash
# round 2, call AC: sampler candidate stage on the GPU -- kernel tests, runner test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sampler.py tests/test_zz_gpu_runner.py -m gpu -x -q -s > gpurun_out/r2ac_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ac_pytest.log
grep -E "\[sampler\]|\[runner\]|passed|failed|Error|error|assert" gpurun_out/r2ac_pytest.log | head -30
