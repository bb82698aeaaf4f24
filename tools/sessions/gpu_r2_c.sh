#!/usr/bin/env bash
# Round-2 GPU session C: occupancy variants with smem trackers, tile launch orders, new pipeline tests.
set -x
mkdir -p gpurun_out
: > gpurun_out/r2c_variants.jsonl
for v in l_mb10 lr_mb9 l_mb9; do
  timeout 300 python tools/variant_bench.py --lib build/variants/libmnv_b200_$v.so --anchor 8 --tag $v >> gpurun_out/r2c_variants.jsonl 2>> gpurun_out/r2c_variants.err
done
for o in morton cols snake; do
  timeout 300 python tools/variant_bench.py --anchor 8 --order $o --tag order_$o >> gpurun_out/r2c_variants.jsonl 2>> gpurun_out/r2c_variants.err
done
cat gpurun_out/r2c_variants.jsonl
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_refine_gpu.py tests/test_interop_gpu.py -x -q 2>&1 | tail -8 | tee gpurun_out/r2c_pytest.log
