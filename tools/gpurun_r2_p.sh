#!/bin/bash
# round 2, call P: delivery-rate ceiling of the whole-token kernel (-DB200_NO_MATH=1: the row loops consume the ring without doing the math)
mkdir -p gpurun_out
timeout 600 python -c "import bench; bench.ensure_model(32)" > gpurun_out/model.log 2>&1
ln -sf /tmp/b200_bench/ggml-model-q4_0.bin /tmp/probe-7b-l32.bin
for v in nomath llama; do
  echo "== $v" >> gpurun_out/r2p_probe.log
  B200_LIB=$PWD/llama.swift_b200/libb200$([ $v = nomath ] && echo _nomath || echo llama).so timeout 300 python tools/probe.py --layers 32 --steps 512 --n-past 8 2>&1 | tail -1 >> gpurun_out/r2p_probe.log
done
cat gpurun_out/r2p_probe.log
B200_LIB=$PWD/llama.swift_b200/libb200_nomath.so timeout 300 python tools/phase_profile.py --layers 8 --pos 264 > gpurun_out/r2p_phase264_nomath.log 2>&1
tail -24 gpurun_out/r2p_phase264_nomath.log
B200_PROF_ATT=1 B200_LIB=$PWD/llama.swift_b200/libb200_profatt.so timeout 300 python tools/phase_profile.py --layers 8 --pos 264 > gpurun_out/r2p_phase264_att.log 2>&1
tail -8 gpurun_out/r2p_phase264_att.log
