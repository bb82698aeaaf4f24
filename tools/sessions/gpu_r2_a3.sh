#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
for v in mb9 mb9nodda; do
timeout 300 python tools/variant_bench.py --lib build/variants/libmnv_b200_$v.so --anchor 8 --tag $v | tee -a gpurun_out/r2a2_variants.jsonl
done
