#!/bin/bash
# round 2, call B: tensor-path (IMMA) row loop -- parity tests, speed, tile-size A/B, phase timeline
mkdir -p gpurun_out
timeout 600 python -c "import bench; bench.ensure_model(32)" > gpurun_out/model.log 2>&1
ln -sf /tmp/b200_bench/ggml-model-q4_0.bin /tmp/probe-7b-l32.bin
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
run() { echo "== $1" >> gpurun_out/r2b_probe.log; shift; env "$@" timeout 300 python tools/probe.py --layers 32 --steps 512 --n-past 8 2>&1 | tail -1 >> gpurun_out/r2b_probe.log; }
run imma_default B200_X=0
run small16 B200_LP_SMALL=16
run qkv8 B200_LP_QKV=8
run qkv8_w13_8 B200_LP_QKV=8 B200_LP_W13=8
run dp4a B200_LIB=$PWD/llama.swift_b200/libb200_base.so
cat gpurun_out/r2b_probe.log
timeout 300 python tools/phase_profile.py --layers 8 --pos 264 > gpurun_out/r2b_phase264.log 2>&1
tail -24 gpurun_out/r2b_phase264.log
