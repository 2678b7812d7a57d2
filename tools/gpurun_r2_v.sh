#!/bin/bash
# round 2, call V: hand-over timeline of the tcgen05 prefill kernel (trace build)
mkdir -p gpurun_out
timeout 300 python tools/tc_trace.py > gpurun_out/r2v_tc_trace.log 2>&1; echo "rc=$?" >> gpurun_out/r2v_tc_trace.log
tail -14 gpurun_out/r2v_tc_trace.log
