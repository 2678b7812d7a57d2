#!/bin/bash
# round 2, call AB: tc kernel v9 (21 warps: TMA polled from the MMA warp; scales prefetched; elected arrives), x16 vs x32 TMEM loads
mkdir -p gpurun_out
for v in 1 0; do
  echo "== LD16=$v"
  B200_LIB=$PWD/llama.swift_b200/libb200_ld16_$v.so timeout 600 python -m pytest tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -1
  B200_LIB=$PWD/llama.swift_b200/libb200_ld16_$v.so timeout 300 python tools/prompt_probe.py --layers 2 --n 256 --reps 3 2>&1 | tail -1
done
