#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -c "import bench; bench.ensure_model(32)" > gpurun_out/model.log 2>&1
ln -sf /tmp/b200_bench/ggml-model-q4_0.bin /tmp/probe-7b-l32.bin
B200_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_token -s 12 -c 1 -f -o gpurun_out/mega_token_kernel python tools/probe.py --layers 32 --steps 8 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 160 -c 60 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 24 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
timeout 300 python tools/phase_profile.py --layers 8 --pos 64 > gpurun_out/phase64.log 2>&1
timeout 300 python tools/phase_profile.py --layers 8 --pos 264 > gpurun_out/phase264.log 2>&1
tail -24 gpurun_out/phase264.log
