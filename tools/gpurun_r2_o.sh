#!/bin/bash
# round 2, call O: mmap loader with two-part files, load timing, Q4 sweep, full suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -4 gpurun_out/r2o_pytest.log
timeout 600 python tools/load_bench.py > gpurun_out/r2o_load.log 2>&1; cat gpurun_out/r2o_load.log
timeout 600 python tools/sweep_q4.py > gpurun_out/r2o_sweep_q4.md 2>&1; cat gpurun_out/r2o_sweep_q4.md
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
