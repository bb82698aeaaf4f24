#!/usr/bin/env bash
# Round-2 GPU session: ncu launch list of the default bench command + --set full captures of the two shipped hot kernels.
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-headless --no-target > gpurun_out/r2n_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_voxels_kernel -s 4 -c 2 -f -o gpurun_out/r2n_traversal python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-headless --no-target > gpurun_out/r2n_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_forward_kernel -s 3 -c 1 -f -o gpurun_out/r2n_mlp python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-headless --no-target > gpurun_out/r2n_ncu3.log 2>&1
ls -la gpurun_out/r2n_*
