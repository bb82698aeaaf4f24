#!/usr/bin/env bash
# Round-2 GPU session I: selection after the two-pass collect, full GPU suite, ncu captures of the shipped kernels.
set -x
mkdir -p gpurun_out
timeout 300 python tools/profile_select.py 2>&1 | tail -5 | tee gpurun_out/r2i_select.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2i_pytest.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2i_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-headless --no-target > gpurun_out/r2i_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_voxels_kernel -s 4 -c 2 -o gpurun_out/r2i_traversal python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-headless --no-target > gpurun_out/r2i_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_forward_kernel -s 3 -c 1 -o gpurun_out/r2i_mlp python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-headless --no-target > gpurun_out/r2i_ncu3.log 2>&1
ls -la gpurun_out/r2i_*
