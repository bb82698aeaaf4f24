#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
for mode in "1 0" "0 1"; do
  set -- $mode
  export MNV_MLP_PAIR=$1 MNV_MLP_PER=$2
  echo "== pair $1 per $2 (0 = by model)" | tee -a gpurun_out/r2v2_mlp_early_pe.log
  timeout 180 python -m pytest tests/test_mlp_gpu.py -x -q 2>&1 | tail -2 | tee -a gpurun_out/r2v2_mlp_early_pe.log
  timeout 120 python tools/mlp_time.py --tag "pair$1_per$2" 2>&1 | tail -1 | tee -a gpurun_out/r2v2_mlp_early_pe.log
  MNV_MLP_DEBUG=1 timeout 120 python tools/mlp_time.py --lib build/variants/libmnv_b200_mlptiming.so --rows 262144 2>&1 | grep "mlp dbg" | tail -2 | tee -a gpurun_out/r2v2_mlp_early_pe.log
done
