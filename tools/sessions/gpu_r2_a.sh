#!/usr/bin/env bash
# Round-2 GPU session A: parity suite, compositor goldens, smoke, traversal variants (anchor grid levels + compile-time
# variants), short bench.
set -x
mkdir -p gpurun_out/golden
python oracle/make_golden.py --only nerf --out gpurun_out/golden 2>&1 | tail -12 | tee gpurun_out/r2a_golden.log
cp gpurun_out/golden/nerf_*.npz tests/golden/ 2>/dev/null
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/r2a_pytest.log | tail -15
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 | tee gpurun_out/r2a_smoke.log
timeout 600 python tools/variant_bench.py --anchor 0,5,6,7,8 --tag base > gpurun_out/r2a_variants.jsonl 2> gpurun_out/r2a_variants.err
for v in lazy regs lazyregs unroll2 lazyu2; do
  timeout 300 python tools/variant_bench.py --lib build/variants/libmnv_b200_$v.so --anchor 0,7 --tag $v >> gpurun_out/r2a_variants.jsonl 2>> gpurun_out/r2a_variants.err
done
cat gpurun_out/r2a_variants.jsonl
timeout 900 python bench.py --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 2500 gpurun_out/r2a_bench.json
