#!/bin/bash
# round 2, call D: arg-max folded into the token kernel, funnel-shift variant; parity tests of the decode loop
mkdir -p gpurun_out
timeout 600 python -c "import bench; bench.ensure_model(32)" > gpurun_out/model.log 2>&1
ln -sf /tmp/b200_bench/ggml-model-q4_0.bin /tmp/probe-7b-l32.bin
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -4 gpurun_out/r2d_pytest.log
run() { echo "== $1" >> gpurun_out/r2d_probe.log; shift; env "$@" timeout 300 python tools/probe.py --layers 32 --steps 512 --n-past 8 2>&1 | tail -2 >> gpurun_out/r2d_probe.log; }
run fold B200_X=0
run shf B200_LIB=$PWD/llama.swift_b200/libb200_shf.so
run base_r1 B200_LIB=$PWD/llama.swift_b200/libb200_base.so
cat gpurun_out/r2d_probe.log
