#!/bin/bash
# round 2, call R: ceilings of the tcgen05 prefill kernel (which stage bounds the ~720-cycle block step)
mkdir -p gpurun_out
for v in llama _tcskip1 _tcskip2 _tcskip3; do
  lib=llama.swift_b200/libb200$v.so; [ $v = llama ] && lib=llama.swift_b200/libb200llama.so
  echo "== $v" >> gpurun_out/r2r_probe.log
  B200_LIB=$PWD/$lib timeout 300 python tools/prompt_probe.py --layers 2 --n 256 --reps 3 2>&1 | tail -1 >> gpurun_out/r2r_probe.log
done
cat gpurun_out/r2r_probe.log
