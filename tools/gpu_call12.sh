#!/bin/bash
mkdir -p gpurun_out
B200_PROF_LN=1 B200_LIB=$PWD/llama.swift_b200/libb200llama_profln.so timeout 300 python tools/phase_profile.py --layers 8 --pos 64 > gpurun_out/phase_ln.log 2>&1; tail -32 gpurun_out/phase_ln.log
