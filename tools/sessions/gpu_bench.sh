#!/usr/bin/env bash
# Dev: GPU test suite + bench line (+ reference arm) + ncu launch list of the bench command.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log | tail -6
python bench.py --steps 100 --warmup 10 > gpurun_out/bench_native.json 2> gpurun_out/bench_native.err; tail -c 3000 gpurun_out/bench_native.json
python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 1500 gpurun_out/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-headless > gpurun_out/b_ncu.log 2>&1
tail -2 gpurun_out/b_ncu.log | cut -c1-300
