#!/usr/bin/env bash
# Round-2 GPU session D: the default bench line at N=1 (with target_4k) + reference arm; wall-clock of each.
set -x
mkdir -p gpurun_out
( time python bench.py > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err ) 2> gpurun_out/r2d_time_n1.txt
tail -c 6000 gpurun_out/r2d_bench_n1.json; tail -5 gpurun_out/r2d_bench_n1.err; cat gpurun_out/r2d_time_n1.txt
( time python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/r2d_bench_ref.json 2> gpurun_out/r2d_bench_ref.err ) 2> gpurun_out/r2d_time_ref.txt
tail -c 1500 gpurun_out/r2d_bench_ref.json; cat gpurun_out/r2d_time_ref.txt
