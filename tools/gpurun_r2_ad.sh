#!/bin/bash
# round 2, call AD (8 x B200): 13B at tp 8 / 4 (BASELINE.json configs[3]) with the fixed multi-part loader, TP tests with the jitter build
mkdir -p gpurun_out
timeout 900 python -c "
import bench
bench.MODEL='13b'; bench.ensure_model(40)" > gpurun_out/model.log 2>&1
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2972$n bench.py --model 13b --gpus $n --steps 128 --warmup 4 --no-replicas > gpurun_out/r2ad_bench_13b_tp$n.json 2> gpurun_out/r2ad_bench_13b_tp$n.err
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2ad_bench_13b_tp$n.json') if l.startswith('{')][-1]); print('13B tp$n', d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity']['ok'], d['parity']['bit_identical_steps'])"
done
B200_LIB=$PWD/llama.swift_b200/libb200_jitter.so timeout 900 python -m pytest tests/test_gpu_tp.py -m gpu -x -q -s > gpurun_out/r2ad_pytest_tp8_jitter.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ad_pytest_tp8_jitter.log
grep -E "tp\]|TP_WORKER|passed|failed|rc=" gpurun_out/r2ad_pytest_tp8_jitter.log | tail -12
