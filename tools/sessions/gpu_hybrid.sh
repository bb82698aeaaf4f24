#!/usr/bin/env bash
# Dev: hybrid (row blocks x cells) multi-GPU mode. usage: gpu_hybrid.sh N
N=$1
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_multigpu_gpu.py -x -q -k "hybrid or across" 2>&1 | tail -4
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" 2>gpurun_out/err_$N.log | grep '^{' ; tail -2 gpurun_out/err_$N.log | cut -c1-300; }
run --steps 60 --warmup 5 --mode hybrid | tee gpurun_out/bench_hybrid_n$N.json | cut -c1-300
