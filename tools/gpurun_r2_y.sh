#!/bin/bash
# round 2, call Y: tc kernel v7 (merged producer warps, 9-instruction unpack, sleeping waits for the roles with slack)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_batch.py -m gpu -x -q > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log
tail -3 gpurun_out/r2y_pytest.log
echo "== epilogue sleeping wait"; timeout 300 python tools/prompt_probe.py --layers 2 --n 256 --reps 3 2>&1 | tail -1
echo "== epilogue polling wait"; B200_LIB=$PWD/llama.swift_b200/libb200_episleep0.so timeout 300 python tools/prompt_probe.py --layers 2 --n 256 --reps 3 2>&1 | tail -1
timeout 300 python tools/tc_trace.py > gpurun_out/r2y_tc_trace.log 2>&1; tail -13 gpurun_out/r2y_tc_trace.log
