#!/bin/bash
# round 2, call W: lean tc kernel (address/counter hoisting, elected MMA issue, sparse-clock waits), hint vs no-hint waits
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_batch.py -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2w_pytest.log
tail -3 gpurun_out/r2w_pytest.log
echo "== hint"; timeout 300 python tools/prompt_probe.py --layers 2 --n 256 --reps 3 2>&1 | tail -1
echo "== nohint"; B200_LIB=$PWD/llama.swift_b200/libb200_nohint.so timeout 300 python tools/prompt_probe.py --layers 2 --n 256 --reps 3 2>&1 | tail -1
timeout 300 python tools/tc_trace.py > gpurun_out/r2w_tc_trace.log 2>&1; tail -13 gpurun_out/r2w_tc_trace.log
