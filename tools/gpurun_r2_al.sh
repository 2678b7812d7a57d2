#!/bin/bash
# round 2, call AL: last check of the final build -- smoke, decode parity, Q4_1, sampler
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sampler.py -m gpu -q -k "q4_1 or vs_oracle or sampler or candidates or ambiguous" 2>&1 | tail -2
