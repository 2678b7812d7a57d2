#!/bin/bash
# round 2, call AJ (2 x B200): long single-process group enqueue
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tp.py -m gpu -q -s -k long_enqueue > gpurun_out/r2aj_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2aj_pytest.log
grep -E "tp\]|passed|failed|Error|rc=" gpurun_out/r2aj_pytest.log | tail -8
