#!/bin/bash
# GPU session 1 (1 x B200): parity tests, A/B of the row-loop / dataflow variants, phase profile, bench, ncu.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
# full-size model once (shared by bench.py and the probes)
timeout 600 python -c "import bench; bench.ensure_model(32)" > gpurun_out/model.log 2>&1
ln -sf /tmp/b200_bench/ggml-model-q4_0.bin /tmp/probe-7b-l32.bin
P="timeout 300 python tools/probe.py --layers 32 --steps 64"
for v in new nopipe v1; do
  lib=llama.swift_b200/libb200llama.so
  [ $v = nopipe ] && lib=llama.swift_b200/libb200llama_nopipe.so
  [ $v = v1 ] && lib=llama.swift_b200/libb200llama_v1.so
  echo "== $v" >> gpurun_out/ab.log
  B200_LIB=$PWD/$lib $P 2>&1 | grep decode | tail -1 >> gpurun_out/ab.log
done
for e in "B200_L2_AHEAD=32" "B200_L2_AHEAD=128" "B200_LP_SMALL=2" "B200_LP_QKV=1" "B200_LP_W13=4" "B200_LP_QKV=4 B200_LP_W13=4" "B200_STAGE_BYTES=24576" "B200_STAGE_BYTES=49152"; do
  echo "== new $e" >> gpurun_out/ab.log
  env $e $P 2>&1 | grep decode | tail -1 >> gpurun_out/ab.log
done
cat gpurun_out/ab.log
timeout 300 python tools/phase_profile.py --layers 8 --pos 64 > gpurun_out/phase.log 2>&1; tail -22 gpurun_out/phase.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json
B200_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_token -s 12 -c 1 -f -o gpurun_out/mega_r1_e python tools/probe.py --layers 32 --steps 8 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 160 -c 60 --csv --log-file gpurun_out/launches_r1_e.csv python bench.py --steps 24 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ls -la gpurun_out
