#!/usr/bin/env bash
# Dev: sharded guided-sampling bench. usage: gpu_guided.sh N
N=$1
mkdir -p gpurun_out
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 16 --mode guided 2>gpurun_out/err_guided_$N.log | grep '^{' | tee gpurun_out/bench_guided_n$N.json
tail -5 gpurun_out/err_guided_$N.log | cut -c1-400
