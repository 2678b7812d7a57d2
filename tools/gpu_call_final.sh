#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python -c "import bench; bench.ensure_model(32)" > gpurun_out/model.log 2>&1
ln -sf /tmp/b200_bench/ggml-model-q4_0.bin /tmp/probe-7b-l32.bin
P="timeout 300 python tools/probe.py --layers 32 --steps 64"
rm -f gpurun_out/ab.log
for v in new nolop3; do
  lib=llama.swift_b200/libb200llama.so
  [ $v != new ] && lib=llama.swift_b200/libb200llama_$v.so
  echo "== $v" >> gpurun_out/ab.log
  B200_LIB=$PWD/$lib $P 2>&1 | grep -E "decode|rror" | tail -1 >> gpurun_out/ab.log
  B200_LIB=$PWD/$lib $P --n-past 256 2>&1 | grep -E "decode|rror" | tail -1 >> gpurun_out/ab.log
done
cat gpurun_out/ab.log
timeout 600 python bench.py > gpurun_out/bench_1.json 2> gpurun_out/bench_1.err
cat gpurun_out/bench_1.json | cut -c1-200; python -c "
import json; d=json.load(open('gpurun_out/bench_1.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])"
