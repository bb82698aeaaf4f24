#!/usr/bin/env bash
# Round-2 GPU session X: with a cheap issue loop (CTA pairs, 4 MMAs per ring stage), which of the concurrent consumers slows the MMAs?
set -x
mkdir -p gpurun_out
for mode in "1 4" "1 2"; do
set -- $mode
export MNV_MLP_PAIR=$1 MNV_MLP_PER=$2
for v in mlptiming mlpd1 mlpd2 mlpd3; do
  MNV_MLP_DEBUG=1 timeout 120 python tools/mlp_time.py --lib build/variants/libmnv_b200_$v.so --rows 262144 --tag $v 2>&1 | grep -E "mlp dbg|rows" | tail -3 | sed "s/^/pair$1 per$2 $v /" | tee -a gpurun_out/r2x2_mlp_diag_lean.log
done
done
