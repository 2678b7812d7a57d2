#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python -c "import bench; bench.ensure_model(32)" > gpurun_out/model.log 2>&1
ln -sf /tmp/b200_bench/ggml-model-q4_0.bin /tmp/probe-7b-l32.bin
P="timeout 300 python tools/probe.py --layers 32 --steps 64"
rm -f gpurun_out/ab.log
echo "== new" >> gpurun_out/ab.log; $P 2>&1 | grep -E "decode|rror" | tail -1 >> gpurun_out/ab.log
echo "== new (n_past 400)" >> gpurun_out/ab.log; $P --n-past 400 2>&1 | grep -E "decode|rror" | tail -1 >> gpurun_out/ab.log
cat gpurun_out/ab.log
timeout 300 python tools/phase_profile.py --layers 8 --pos 64 > gpurun_out/phase.log 2>&1; tail -24 gpurun_out/phase.log
timeout 300 python tools/phase_profile.py --layers 8 --pos 400 > gpurun_out/phase400.log 2>&1; tail -24 gpurun_out/phase400.log | grep -E "attention|per layer|arrival"
