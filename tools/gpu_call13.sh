#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python -c "import bench; bench.ensure_model(32)" > gpurun_out/model.log 2>&1
ln -sf /tmp/b200_bench/ggml-model-q4_0.bin /tmp/probe-7b-l32.bin
P="timeout 300 python tools/probe.py --layers 32 --steps 64"
rm -f gpurun_out/ab.log
for v in new nostage; do
  lib=llama.swift_b200/libb200llama.so
  [ $v != new ] && lib=llama.swift_b200/libb200llama_$v.so
  echo "== $v" >> gpurun_out/ab.log
  B200_LIB=$PWD/$lib $P 2>&1 | grep -E "decode|rror" | tail -1 >> gpurun_out/ab.log
  B200_LIB=$PWD/$lib $P --n-past 256 2>&1 | grep -E "decode|rror" | tail -1 >> gpurun_out/ab.log
done
cat gpurun_out/ab.log
timeout 300 python tools/phase_profile.py --layers 8 --pos 64 > gpurun_out/phase.log 2>&1; tail -22 gpurun_out/phase.log
