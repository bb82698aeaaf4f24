#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
python tools/variant_bench.py --anchor 8 --tag packraw_default | tee -a gpurun_out/r2a2_variants.jsonl
timeout 300 python tools/variant_bench.py --lib build/variants/libmnv_b200_mb7.so --anchor 8 --tag mb7 | tee -a gpurun_out/r2a2_variants.jsonl
