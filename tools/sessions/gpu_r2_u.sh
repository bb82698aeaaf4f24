#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
for v in mlpt4 mlpt7; do
MNV_MLP_DEBUG=1 timeout 300 python tools/mlp_time.py --lib build/variants/libmnv_b200_$v.so --rows 262144 --tag $v 2>&1 | grep -E "mlp dbg|rows" | tail -2 | sed "s/^/$v /" | tee -a gpurun_out/r2u_mlp_diag.log
done
cd tools/micro && nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o umma_bench umma_bench.cu && ./umma_bench | tee ../../gpurun_out/r2u_umma_bench.log
