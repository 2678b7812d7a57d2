#!/bin/bash
# round 2, call A: baseline of the round-1 kernel on today's box + idle-L2-prefetch A/B + LayerNorm-mark timeline
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
timeout 600 python -c "import bench; bench.ensure_model(32)" > gpurun_out/model.log 2>&1
ln -sf /tmp/b200_bench/ggml-model-q4_0.bin /tmp/probe-7b-l32.bin
for v in base pf4 pf16; do
  lib=llama.swift_b200/libb200_$v.so; true
  echo "== $v" >> gpurun_out/r2a_probe.log
  B200_LIB=$PWD/$lib timeout 300 python tools/probe.py --layers 32 --steps 512 --n-past 8 >> gpurun_out/r2a_probe.log 2>&1
done
B200_PROF_LN=1 B200_LIB=$PWD/llama.swift_b200/libb200_ln.so timeout 300 python tools/phase_profile.py --layers 8 --pos 264 > gpurun_out/r2a_phase_ln264.log 2>&1
B200_LIB=$PWD/llama.swift_b200/libb200_base.so timeout 300 python tools/phase_profile.py --layers 8 --pos 264 > gpurun_out/r2a_phase264.log 2>&1
grep -E "==|decode" gpurun_out/r2a_probe.log
tail -30 gpurun_out/r2a_phase_ln264.log
