#!/usr/bin/env bash
# Round-2 GPU session J: the restructured MLP issuer (weight chunk shared by both tiles).  Everything under timeouts:
# a dead-locked kernel must not hang the box.
set -x
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_mlp_gpu.py -x -q 2>&1 | tail -12 | tee gpurun_out/r2j_pytest_mlp.log | cut -c1-220
rc=${PIPESTATUS[0]}
if [ "$rc" != "0" ]; then echo "MLP tests failed or timed out (rc=$rc)"; exit 0; fi
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/r2j_mlp.log
import sys, json, numpy as np, torch
sys.path.insert(0, ".")
import bench, mega_nerf_viewer_b200 as mnv
dev = torch.device("cuda", 0)
print(json.dumps(bench.mlp_section(mnv, torch, dev)))
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
PY
timeout 600 python -m pytest tests/test_pipeline_gpu.py tests/test_viewer_gpu.py tests/test_group_gpu.py -x -q 2>&1 | tail -5 | tee gpurun_out/r2j_pytest2.log | cut -c1-200
timeout 300 python tools/profile_select.py --width 1920 --height 1080 2>&1 | tail -4 | tee gpurun_out/r2j_select_1080p.log
timeout 300 python tools/profile_select.py 2>&1 | tail -4 | tee gpurun_out/r2j_select_4k.log
