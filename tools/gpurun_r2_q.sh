#!/bin/bash
# round 2, call Q: paced weight streaming experiment + attention sub-phase stamps
mkdir -p gpurun_out
timeout 600 python -c "import bench; bench.ensure_model(32)" > gpurun_out/model.log 2>&1
ln -sf /tmp/b200_bench/ggml-model-q4_0.bin /tmp/probe-7b-l32.bin
# per-SM rate: 1 byte per X ps; chip rate = 148 / X TB/s: 49 ps -> 3.0 TB/s, 42 -> 3.5, 37 -> 4.0, 30 -> 4.9, 25 -> 5.9
for x in 0 49 42 37 30 25; do
  echo "== pace $x ps/byte" >> gpurun_out/r2q_probe.log
  B200_PACE_PS_PER_BYTE=$x timeout 300 python tools/probe.py --layers 32 --steps 512 --n-past 8 2>&1 | tail -1 >> gpurun_out/r2q_probe.log
done
cat gpurun_out/r2q_probe.log
B200_PROF_ATT=1 B200_LIB=$PWD/llama.swift_b200/libb200_profatt.so timeout 300 python tools/phase_profile.py --layers 8 --pos 264 > gpurun_out/r2q_phase264_att.log 2>&1
tail -8 gpurun_out/r2q_phase264_att.log
