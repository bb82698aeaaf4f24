#!/usr/bin/env bash
# Dev: one gpurun call = GPU test suite + MLP timing + headless driver on the bench tree.
set -x
mkdir -p gpurun_out
MNV_GOLDEN_OUT=gpurun_out/golden python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
python tools/mlp_check.py > gpurun_out/mlp_check.log 2>&1; tail -5 gpurun_out/mlp_check.log
ROWS=2097152 python tools/mlp_check.py > gpurun_out/mlp_check_2m.log 2>&1; tail -3 gpurun_out/mlp_check_2m.log
