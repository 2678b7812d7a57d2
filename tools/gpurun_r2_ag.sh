#!/bin/bash
# round 2, call AG: final validation -- full GPU test suite, bench (decode, prefill), smoke, runner sampler A/B, ncu launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2ag_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ag_pytest_gpu.log
tail -4 gpurun_out/r2ag_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2ag_bench.json 2> gpurun_out/r2ag_bench.err; tail -c 1500 gpurun_out/r2ag_bench.json
timeout 600 python bench.py --mode prefill > gpurun_out/r2ag_bench_prefill.json 2> gpurun_out/r2ag_bench_prefill.err; tail -c 1200 gpurun_out/r2ag_bench_prefill.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python tools/run_bench.py --n-predict 256 > gpurun_out/r2ag_run_bench.log 2>&1; cat gpurun_out/r2ag_run_bench.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2ag_launches_bench.csv python bench.py --steps 8 --warmup 3 --no-parity > gpurun_out/r2ag_ncu_bench.log 2>&1; tail -2 gpurun_out/r2ag_ncu_bench.log | cut -c1-300
