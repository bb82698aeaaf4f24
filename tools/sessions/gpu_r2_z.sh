#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
for mode in "0 1" "1 4"; do
  set -- $mode
  export MNV_MLP_PAIR=$1 MNV_MLP_PER=$2
for v in mlptiming mlpd2; do
  MNV_MLP_DEBUG=1 timeout 120 python tools/mlp_time.py --lib build/variants/libmnv_b200_$v.so --rows 262144 --tag $v 2>&1 | grep -E "mlp dbg" | tail -2 | sed "s/^/pair$1 per$2 $v /" | tee -a gpurun_out/r2z_mlp_epilogue_timing.log
done
done
