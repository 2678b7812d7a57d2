#!/bin/bash
# 2 x B200: tensor-parallel group tests (single-process group + one-process-per-GPU IPC), then bench at N=2 and N=1.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_tp.py -m gpu -x -q -s > gpurun_out/pytest_tp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tp.log
tail -30 gpurun_out/pytest_tp.log
timeout 600 python -c "import bench; bench.ensure_model(32)" > gpurun_out/model.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 128 --warmup 4 > gpurun_out/bench_tp2.json 2> gpurun_out/bench_tp2.err
cat gpurun_out/bench_tp2.json; tail -5 gpurun_out/bench_tp2.err
timeout 600 python bench.py --steps 128 --warmup 4 --no-cpu-baseline > gpurun_out/bench_1.json 2> gpurun_out/bench_1.err
cat gpurun_out/bench_1.json
