#!/bin/bash
# round 2, call AA: source-level ncu capture of tc kernel v8
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:q4_gemm_tc -s 4 -c 2 -o gpurun_out/r2aa_tc -f \
  python tools/prompt_probe.py --layers 2 --n 256 --reps 1 > gpurun_out/r2aa_ncu.log 2>&1
tail -3 gpurun_out/r2aa_ncu.log
ls -la gpurun_out/r2aa_tc.ncu-rep
