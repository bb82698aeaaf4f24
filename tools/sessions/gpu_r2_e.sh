#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -60 | tee gpurun_out/r2e_pytest.log | cut -c1-200
