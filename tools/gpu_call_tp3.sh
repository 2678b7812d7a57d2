#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tp.py -m gpu -x -q -s > gpurun_out/pytest_tp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tp.log
grep -E "tp\]|TP_WORKER|passed|failed|rc=|rror" gpurun_out/pytest_tp.log | tail -12
timeout 600 python -c "import bench; bench.ensure_model(32)" > gpurun_out/model.log 2>&1
for v in new llh; do
  lib=llama.swift_b200/libb200llama.so
  [ $v != new ] && lib=llama.swift_b200/libb200llama_$v.so
  B200_LIB=$PWD/$lib timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 128 --warmup 4 > gpurun_out/bench_tp2_$v.json 2> gpurun_out/bench_tp2_$v.err
  python -c "
import json,sys
for l in open('gpurun_out/bench_tp2_$v.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$v', d['value'], d['ms_per_step'], d['e2e']['value'])
"
done
timeout 300 python tools/sweep_q4.py > gpurun_out/sweep_q4.md 2>&1; cat gpurun_out/sweep_q4.md
