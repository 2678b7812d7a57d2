#!/bin/bash
# round 2, call AI (2 x B200): tensor-parallel tests and the 2-GPU bench with the final build (eval path refactored for the GPU sampler)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tp.py -m gpu -q -s > gpurun_out/r2ai_pytest_tp2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ai_pytest_tp2.log
grep -E "tp\]|TP_WORKER|passed|failed|rc=" gpurun_out/r2ai_pytest_tp2.log | tail -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 128 --warmup 4 --no-replicas > gpurun_out/r2ai_bench_7b_tp2.json 2> gpurun_out/r2ai_bench_7b_tp2.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2ai_bench_7b_tp2.json') if l.startswith('{')][-1]); print('7B tp2', d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity']['ok'], d['parity']['bit_identical_steps'])"
