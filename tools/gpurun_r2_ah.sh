#!/bin/bash
# round 2, call AH: compute-sanitizer memcheck + racecheck over every path incl. the round's new kernels (tools/sanitize_probe.py)
mkdir -p gpurun_out
timeout 100 python tools/sanitize_probe.py 2>&1 | tail -2
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_probe.py > gpurun_out/r2ah_memcheck.log 2>&1; tail -5 gpurun_out/r2ah_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --print-limit 30 python tools/sanitize_probe.py > gpurun_out/r2ah_racecheck.log 2>&1; grep -E "RACECHECK SUMMARY|sanitize probe|Error:|Warning:" gpurun_out/r2ah_racecheck.log | sort | uniq -c | sort -rn | head -12
grep -A3 -m3 "hazard" gpurun_out/r2ah_racecheck.log | cut -c1-220 | head -30
