#!/bin/bash
# round 2, call K: bench.py (decode, 512 steps) + bench.py --mode prefill + ncu DRAM traffic of the decode kernel + full GPU suite
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2k_bench_decode.json 2> gpurun_out/r2k_bench_decode.err
python -c "
import json; d=json.load(open('gpurun_out/r2k_bench_decode.json')); print('decode', d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity'], d['cpu_baseline']['value'])"
timeout 900 python bench.py --mode prefill --steps 3 --warmup 3 > gpurun_out/r2k_bench_prefill.json 2> gpurun_out/r2k_bench_prefill.err
python -c "
import json; d=json.load(open('gpurun_out/r2k_bench_prefill.json')); print('prefill', d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity'])"
tail -2 gpurun_out/r2k_bench_prefill.err
ln -sf /tmp/b200_bench/ggml-model-q4_0.bin /tmp/probe-7b-l32.bin
B200_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_token -s 12 -c 1 -f -o gpurun_out/r2k_mega python tools/probe.py --layers 32 --steps 8 > gpurun_out/r2k_ncu_full.log 2>&1
tail -2 gpurun_out/r2k_ncu_full.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 160 -c 60 --csv --log-file gpurun_out/r2k_launches_bench.csv python bench.py --steps 24 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2k_ncu_launch.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -4 gpurun_out/r2k_pytest.log
timeout 300 python tools/phase_profile.py --layers 8 --pos 264 > gpurun_out/r2k_phase264.log 2>&1
