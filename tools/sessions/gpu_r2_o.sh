#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
MNV_B200_LIB=$PWD/build/variants/libmnv_b200_mlptiming.so MNV_MLP_DEBUG=1 timeout 300 python - <<'PY' 2>&1 | tail -8 | tee gpurun_out/r2o_mlp_timing.log
import sys, numpy as np, torch
sys.path.insert(0, ".")
import mega_nerf_viewer_b200 as mnv
dev = torch.device("cuda", 0)
model = mnv.MlpModel([mnv.synth.make_mlp_weights(seed=3)], device=0)
rows = 262144
x = torch.rand((rows, model.in_dim), device=dev) * 2 - 1; x[:, -1] = 0
out = torch.empty((rows, model.out_dim + 1), device=dev)
for _ in range(4): model.forward(x, out=out)
torch.cuda.synchronize()
PY
timeout 300 python tools/profile_select.py --width 1920 --height 1080 2>&1 | tail -5 | tee gpurun_out/r2o_select_1080p.log
