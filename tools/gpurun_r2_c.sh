#!/bin/bash
# round 2, call C: pipe micro-benchmark, full GPU test suite (incl. full-size 7B / 13B parity), bench.py with the parity block
mkdir -p gpurun_out
./tools/microbench/pipe_rates > gpurun_out/r2c_pipe_rates.txt 2>&1
cat gpurun_out/r2c_pipe_rates.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
grep -E "passed|failed|rc=|parity-full|FAILED|Error" gpurun_out/r2c_pytest.log | tail -8
timeout 600 python bench.py --steps 128 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
tail -c 1500 gpurun_out/r2c_bench.json; tail -3 gpurun_out/r2c_bench.err
