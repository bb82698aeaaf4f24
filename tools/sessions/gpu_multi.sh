#!/usr/bin/env bash
# Dev: multi-GPU checks. usage: gpu_multi.sh N
N=$1
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -4
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" 2>gpurun_out/err_$N.log | grep '^{' ; tail -3 gpurun_out/err_$N.log | cut -c1-300; }
run --steps 60 --warmup 5 --no-cpu-baseline | tee gpurun_out/bench_tiles_n$N.json | cut -c1-400
run --steps 60 --warmup 5 --mode split | tee gpurun_out/bench_split_n$N.json | cut -c1-600
