#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_tp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tp.log
grep -E "tp\]|TP_WORKER|passed|failed|rc=|FAILED|13B" gpurun_out/pytest_tp.log | tail -30
