#!/bin/bash
# 8 x B200: tensor-parallel groups of 2 / 4 / 8 (7B-width and 13B-geometry models), then the bench at N=8 and N=4.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tp.py -m gpu -q -s > gpurun_out/pytest_tp8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tp8.log
grep -E "tp\]|TP_WORKER|passed|failed|rc=|FAILED" gpurun_out/pytest_tp8.log | tail -20
timeout 600 python -c "import bench; bench.ensure_model(32)" > gpurun_out/model.log 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --steps 128 --warmup 4 > gpurun_out/bench_tp8.json 2> gpurun_out/bench_tp8.err
grep -E "^\{" gpurun_out/bench_tp8.json | cut -c1-260; tail -3 gpurun_out/bench_tp8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 4 --steps 128 --warmup 4 > gpurun_out/bench_tp4.json 2> gpurun_out/bench_tp4.err
grep -E "^\{" gpurun_out/bench_tp4.json | cut -c1-260; tail -3 gpurun_out/bench_tp4.err
