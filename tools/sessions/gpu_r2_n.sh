#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/r2n_mlp_shipped.log
import sys, json, numpy as np, torch
sys.path.insert(0, ".")
import mega_nerf_viewer_b200 as mnv
dev = torch.device("cuda", 0)
model = mnv.MlpModel([mnv.synth.make_mlp_weights(seed=3)], device=0)
for rows in (4096, 32768, 262144, 2097152):
    x = torch.rand((rows, model.in_dim), device=dev) * 2 - 1; x[:, -1] = 0
    out = torch.empty((rows, model.out_dim + 1), device=dev)
    for _ in range(3): model.forward(x, out=out)
    torch.cuda.synchronize(); ms = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); model.forward(x, out=out); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    t = float(np.mean(ms)); print(rows, "rows", round(t, 4), "ms", round(rows * model.flops_per_row / t / 1e9, 1), "TFLOP/s")
# the reference's evaluation mode: torch fp16 autocast of the same module (cuda_renderer.cpp:188-193)
sys.path.insert(0, "tests")
from mlp_reference import MegaNerfMLP
torch.manual_seed(3)
ref = MegaNerfMLP().cuda().eval()
rows = 262144
x = torch.rand((rows, 4), device=dev) * 2 - 1; x[:, -1] = 0
with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
    for _ in range(3): ref(x)
    torch.cuda.synchronize(); ms = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ref(x); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
print("torch eager fp16 autocast, 262144 rows:", round(float(np.mean(ms)), 4), "ms")
PY
