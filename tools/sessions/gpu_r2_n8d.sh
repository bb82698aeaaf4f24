#!/usr/bin/env bash
# Config 3 / 5 at the scale they exist for: 4K, Mill-19-scale tree, one spatial cell + sub-MLP per GPU on 8 GPUs, with in-run parity.
set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --mode split --workload mill19 --steps 24 > gpurun_out/r2_bench_split_mill19_n8.json 2> gpurun_out/r2_bench_split_mill19_n8.err
tail -c 1200 gpurun_out/r2_bench_split_mill19_n8.json; tail -2 gpurun_out/r2_bench_split_mill19_n8.err | cut -c1-300
MNV_MLP_PAIR=0 MNV_MLP_PER=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --mode guided --workload mill19 --width 3840 --height 2160 --steps 6 > gpurun_out/r2_bench_guided_mill19_4k_n8.json 2> gpurun_out/r2_bench_guided_mill19_4k_n8.err
tail -c 1500 gpurun_out/r2_bench_guided_mill19_4k_n8.json; tail -2 gpurun_out/r2_bench_guided_mill19_4k_n8.err | cut -c1-300
