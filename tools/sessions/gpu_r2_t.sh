#!/usr/bin/env bash
# Round-2 GPU session T: MLP issue loop with the prefetched poll: tests, timing, issuer counters; traversal default after DDA sign selection.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mlp_gpu.py -x -q 2>&1 | tail -3 | tee gpurun_out/r2t_pytest_mlp.log
timeout 300 python tools/mlp_time.py --tag prefetch 2>&1 | tail -1 | tee gpurun_out/r2t_mlp_time.log
MNV_MLP_DEBUG=1 timeout 300 python tools/mlp_time.py --lib build/variants/libmnv_b200_mlpt_pf.so --rows 262144 2>&1 | grep "mlp dbg" | tail -2 | tee gpurun_out/r2t_mlp_issuer.log
timeout 300 python tools/variant_bench.py --anchor 8 --tag dda_default | tee gpurun_out/r2t_variants.jsonl
