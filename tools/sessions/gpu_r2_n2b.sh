#!/usr/bin/env bash
# Round-2 second pass on two GPUs: full GPU suite (incl. the multi-process tests), default bench, split / sharded-guided with in-run parity.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2b_pytest_2gpu.log | cut -c1-200
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err ) 2> gpurun_out/r2b_time_n2.txt
tail -2 gpurun_out/r2b_bench_n2.err | cut -c1-300; grep real gpurun_out/r2b_time_n2.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --mode split --workload mill19 --steps 24 > gpurun_out/r2b_bench_split_mill19_n2.json 2> gpurun_out/r2b_bench_split_mill19_n2.err
tail -c 900 gpurun_out/r2b_bench_split_mill19_n2.json; tail -3 gpurun_out/r2b_bench_split_mill19_n2.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --mode guided --workload mill19 --width 1920 --height 1080 --steps 8 > gpurun_out/r2b_bench_guided_mill19_n2.json 2> gpurun_out/r2b_bench_guided_mill19_n2.err
tail -c 900 gpurun_out/r2b_bench_guided_mill19_n2.json; tail -3 gpurun_out/r2b_bench_guided_mill19_n2.err | cut -c1-300
