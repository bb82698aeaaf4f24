#!/usr/bin/env bash
# Dev: Mill-19-scale 4K frame on N GPUs: image tiles, then refinement on. usage: gpu_mill19.sh N
N=$1
mkdir -p gpurun_out
free -g | head -2
avail=$(free -g | awk '/^Mem:/ {print $7}')
if [ "$avail" -lt $((N * 25 + 40)) ]; then echo "not enough host memory ($avail GB)"; exit 0; fi
run() { timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload mill19 --max-nodes 20000000 "$@" 2>gpurun_out/err_mill19_$N.log | grep '^{'; tail -3 gpurun_out/err_mill19_$N.log | cut -c1-300; }
time run --steps 20 --warmup 3 --no-cpu-baseline --no-headless | tee gpurun_out/bench_mill19_tiles_n$N.json | cut -c1-500
export MNV_BENCH_REUSE_TREE=1
time run --steps 24 --mode refine | tee gpurun_out/bench_mill19_refine_n$N.json | cut -c1-500
rm -rf /dev/shm/mnv_bench_tree
