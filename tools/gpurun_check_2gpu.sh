#!/bin/bash
# 2 x B200: all GPU tests (incl. tensor-parallel groups), then bench at N=2 and N=1.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_tp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tp.log
grep -E "tp\]|TP_WORKER|passed|failed|rc=" gpurun_out/pytest_tp.log | tail -12
timeout 600 python -c "import bench; bench.ensure_model(32)" > gpurun_out/model.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 > gpurun_out/bench_tp2.json 2> gpurun_out/bench_tp2.err
cat gpurun_out/bench_tp2.json; tail -3 gpurun_out/bench_tp2.err
timeout 600 python bench.py > gpurun_out/bench_1.json 2> gpurun_out/bench_1.err
cat gpurun_out/bench_1.json; tail -3 gpurun_out/bench_1.err
