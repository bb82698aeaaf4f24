#!/usr/bin/env bash
# Round-2 GPU session B: occupancy / tile variants, Mill-19 4K anchor levels, ncu full capture of the shipped kernel.
set -x
mkdir -p gpurun_out
timeout 300 python tools/variant_bench.py --anchor 7,8 --tag default_lr > gpurun_out/r2b_variants.jsonl 2> gpurun_out/r2b_variants.err
for v in lr_mb10 lr_mb12 lr_th16; do
  timeout 300 python tools/variant_bench.py --lib build/variants/libmnv_b200_$v.so --anchor 7 --tag $v >> gpurun_out/r2b_variants.jsonl 2>> gpurun_out/r2b_variants.err
done
timeout 900 python tools/variant_bench.py --mill19 --anchor 0,7,8 --steps 24 --tag mill19 >> gpurun_out/r2b_variants.jsonl 2>> gpurun_out/r2b_variants.err
cat gpurun_out/r2b_variants.jsonl
# ncu: launch list of the bench command + one full capture of the shipped tracking kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-headless > gpurun_out/r2b_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_voxels_kernel -s 4 -c 2 -o gpurun_out/r2b_traversal python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-headless > gpurun_out/r2b_ncu2.log 2>&1
ls -la gpurun_out/r2b_*
