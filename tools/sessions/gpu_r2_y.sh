#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
(cd tools/micro && timeout 60 ./tmem_ld_bench) | tee gpurun_out/r2y_tmem_ld_bench.log
export MNV_MLP_PAIR=1 MNV_MLP_PER=4
for v in mlpd3 mlpd6; do
  MNV_MLP_DEBUG=1 timeout 120 python tools/mlp_time.py --lib build/variants/libmnv_b200_$v.so --rows 262144 --tag $v 2>&1 | grep -E "mlp dbg|rows" | tail -2 | sed "s/^/$v /" | tee -a gpurun_out/r2x_mlp_diag_pair4.log
done
