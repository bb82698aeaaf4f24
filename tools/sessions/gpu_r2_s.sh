#!/usr/bin/env bash
# Round-2 GPU session S: (1) MLP issuer counters with the epilogue's smem stores / the TMA writes removed (diagnostic
# builds, wrong results, timing only); (2) traversal variants: warps per CTA, packed priorities, DDA sign selection.
set -x
mkdir -p gpurun_out
for v in mlpt0 mlpt1 mlpt2 mlpt3; do
MNV_B200_LIB=$PWD/build/variants/libmnv_b200_$v.so MNV_MLP_DEBUG=1 timeout 300 python - <<'PY' 2>&1 | tail -3 | sed "s/^/$v /" | tee -a gpurun_out/r2s_mlp_diag.log
import sys, numpy as np, torch
sys.path.insert(0, ".")
import mega_nerf_viewer_b200 as mnv
dev = torch.device("cuda", 0)
model = mnv.MlpModel([mnv.synth.make_mlp_weights(seed=3)], device=0)
rows = 262144
x = torch.rand((rows, model.in_dim), device=dev) * 2 - 1; x[:, -1] = 0
out = torch.empty((rows, model.out_dim + 1), device=dev)
for _ in range(3): model.forward(x, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
PY
done
python tools/variant_bench.py --anchor 8 --tag shipped | tee -a gpurun_out/r2s_variants.jsonl
for v in w1 w2 dda pack w1pack w2pack w1packdda w2dda; do
  timeout 300 python tools/variant_bench.py --lib build/variants/libmnv_b200_$v.so --anchor 8 --tag $v | tee -a gpurun_out/r2s_variants.jsonl
done
