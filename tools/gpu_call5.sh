#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -c "import bench; bench.ensure_model(32)" > gpurun_out/model.log 2>&1
ln -sf /tmp/b200_bench/ggml-model-q4_0.bin /tmp/probe-7b-l32.bin
P="timeout 300 python tools/probe.py --layers 32 --steps 64"
rm -f gpurun_out/ab.log
echo "== new" >> gpurun_out/ab.log; $P 2>&1 | grep -E "decode|rror" | tail -1 >> gpurun_out/ab.log
for e in "B200_L2_AHEAD=1" "B200_L2_AHEAD=2" "B200_L2_AHEAD=4" "B200_L2_AHEAD=6" "B200_L2_AHEAD=8" "B200_L2_AHEAD=12" "B200_L2_AHEAD=16" "B200_L2_AHEAD=4 B200_STAGE_BYTES=32768" "B200_L2_AHEAD=8 B200_STAGE_BYTES=32768" "B200_L2_AHEAD=16 B200_STAGE_BYTES=32768" "B200_LP_W13=1"; do
  echo "== new $e" >> gpurun_out/ab.log
  env $e $P 2>&1 | grep -E "decode|rror" | tail -1 >> gpurun_out/ab.log
done
cat gpurun_out/ab.log
B200_L2_AHEAD=8 timeout 300 python tools/phase_profile.py --layers 8 --pos 64 > gpurun_out/phase_l2a8.log 2>&1; tail -20 gpurun_out/phase_l2a8.log
