#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_group_gpu.py tests/test_viewer_gpu.py tests/test_pipeline_gpu.py -x -q 2>&1 | tail -4 | tee gpurun_out/r2q_pytest.log | cut -c1-200
