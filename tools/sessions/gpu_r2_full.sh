#!/usr/bin/env bash
# Round-2 GPU session: full GPU suite + smoke + default bench of the current tree.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2f_pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2f_smoke.log
( time timeout 900 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err ) 2> gpurun_out/r2f_time_n1.txt
tail -c 1500 gpurun_out/r2f_bench_n1.json; tail -3 gpurun_out/r2f_bench_n1.err | cut -c1-300
