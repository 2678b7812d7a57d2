#!/bin/bash
# round 2, call L: 8-threads-per-row loop for wo / w2 (parity + speed A/B), launch list of the bench command, prefill bench
mkdir -p gpurun_out
timeout 600 python -c "import bench; bench.ensure_model(32)" > gpurun_out/model.log 2>&1
ln -sf /tmp/b200_bench/ggml-model-q4_0.bin /tmp/probe-7b-l32.bin
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_models.py -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -4 gpurun_out/r2l_pytest.log
run() { echo "== $1" >> gpurun_out/r2l_probe.log; shift; env "$@" timeout 300 python tools/probe.py --layers 32 --steps 512 --n-past 8 2>&1 | tail -2 >> gpurun_out/r2l_probe.log; }
run half_rows B200_X=0
run four_threads B200_HALF_ROWS=0
cat gpurun_out/r2l_probe.log
timeout 300 python tools/phase_profile.py --layers 8 --pos 264 > gpurun_out/r2l_phase264.log 2>&1
grep -E "rows|per layer" gpurun_out/r2l_phase264.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 60 --csv --log-file gpurun_out/r2l_launches_bench.csv python bench.py --steps 24 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2l_ncu_launch.log 2>&1
timeout 900 python bench.py --mode prefill --steps 3 --warmup 3 > gpurun_out/r2l_bench_prefill.json 2> gpurun_out/r2l_bench_prefill.err
python -c "
import json; d=json.load(open('gpurun_out/r2l_bench_prefill.json')); print('prefill', d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity'])"
tail -2 gpurun_out/r2l_bench_prefill.err
