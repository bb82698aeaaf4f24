#!/usr/bin/env bash
# Round-2 GPU session V: MLP kernel modes (CTA pairs with cta_group::2, MMAs per ring stage): tests, timing, issuer counters
set -x
mkdir -p gpurun_out
for mode in "0 1" "0 2" "1 1" "1 2" "1 4"; do
  set -- $mode
  export MNV_MLP_PAIR=$1 MNV_MLP_PER=$2
  echo "== pair $1 per $2" | tee -a gpurun_out/r2v_mlp_modes.log
  timeout 180 python -m pytest tests/test_mlp_gpu.py -x -q 2>&1 | tail -2 | tee -a gpurun_out/r2v_mlp_modes.log
  timeout 120 python tools/mlp_time.py --tag "pair$1_per$2" 2>&1 | tail -1 | tee -a gpurun_out/r2v_mlp_modes.log
  MNV_MLP_DEBUG=1 timeout 120 python tools/mlp_time.py --lib build/variants/libmnv_b200_mlptiming.so --rows 262144 2>&1 | grep "mlp dbg" | tail -1 | tee -a gpurun_out/r2v_mlp_modes.log
done
