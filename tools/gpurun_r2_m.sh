#!/bin/bash
# round 2, call M (2 x B200): TP tests, TP2 phase timeline, bench 7B at N=2 (parity block), 13B at N=1 and N=2, soak + jitter soak
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2m_topo.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_tp.py tests/test_gpu_soak.py tests/test_gpu_batch.py -m gpu -x -q -s > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
grep -E "tp\]|TP_WORKER|soak|passed|failed|rc=|FAILED" gpurun_out/r2m_pytest.log | tail -12
timeout 300 python tools/phase_profile.py --layers 8 --pos 264 --tp 2 > gpurun_out/r2m_phase264_tp2.log 2>&1
tail -24 gpurun_out/r2m_phase264_tp2.log
timeout 600 python -c "import bench; bench.ensure_model(32)" > gpurun_out/model.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 > gpurun_out/r2m_bench_tp2.json 2> gpurun_out/r2m_bench_tp2.err
python -c "
import json; d=json.load(open('gpurun_out/r2m_bench_tp2.json')); print('7B tp2', d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity'])"; tail -2 gpurun_out/r2m_bench_tp2.err
timeout 900 python bench.py --model 13b --steps 128 --warmup 4 > gpurun_out/r2m_bench_13b_n1.json 2> gpurun_out/r2m_bench_13b_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r2m_bench_13b_n1.json')); print('13B n1', d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity'], d['cpu_baseline']['value'])"; tail -2 gpurun_out/r2m_bench_13b_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --model 13b --gpus 2 --steps 128 --warmup 4 > gpurun_out/r2m_bench_13b_tp2.json 2> gpurun_out/r2m_bench_13b_tp2.err
python -c "
import json; d=json.load(open('gpurun_out/r2m_bench_13b_tp2.json')); print('13B tp2', d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity'])"; tail -2 gpurun_out/r2m_bench_13b_tp2.err
B200_LIB=$PWD/llama.swift_b200/libb200_jitter.so timeout 900 python -m pytest tests/test_gpu_tp.py -m gpu -x -q -k "single_process" > gpurun_out/r2m_pytest_jitter.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest_jitter.log
tail -3 gpurun_out/r2m_pytest_jitter.log
