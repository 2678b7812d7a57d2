#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=|cache\]|FAILED|Error" gpurun_out/pytest_gpu.log | tail -6
timeout 600 python bench.py > gpurun_out/bench_1.json 2> gpurun_out/bench_1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_1.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'], d['cpu_baseline']['value'])"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
