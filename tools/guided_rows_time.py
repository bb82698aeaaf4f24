"""Dev: the MLP share of a guided-sampling row block on REAL emitted rows (Mill-19-scale tree, 960x540, the lower half of
the frame, 2 sub-modules on a (1, 2) grid) — pair vs single-CTA MLP mode, one process, one GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mega_nerf_viewer_b200 as mnv
import bench
W, H = 960, 540
tree = mnv.synth.make_tree(**bench.mill19_params(24_000_000))
dt = mnv.DeviceTree(tree, device=0)
grid, mn, rng = (1, 2), (-1, -1, -1), (2, 2, 2)
model = mnv.MlpModel([mnv.synth.make_mlp_weights(seed=11 + i) for i in range(2)], grid_dim=grid, min_position=mn, max_position=(1, 1, 1), device=0)
gopt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8], use_guided_sampling=True, appearance_embedding=0)
cam = mnv.synth.default_camera(W, H, pose=3, n_poses=16)
wc = mnv.multigpu.window_camera(cam, 272, 268)
g = dt.guided_samples(wc, gopt, grid, mn, rng, capacity_rows=W * H * 12)
total = g["total"]
cl, rows = g["cluster"][:total], g["rows"][:total]
vals = torch.zeros((total, tree.data_dim + 1), device="cuda")
print("rows", total, "per sub-module", np.bincount(cl.cpu().numpy().astype(np.int64), minlength=2).tolist(), flush=True)
for _ in range(3):
    model.query_submodules(cl, rows, vals)
torch.cuda.synchronize()
ms = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); model.query_submodules(cl, rows, vals); e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
ms.sort()
print(f"pair={os.environ.get('MNV_MLP_PAIR', '1')}: query_submodules on {total} emitted rows: {ms[len(ms) // 2]:.3f} ms "
      f"({total * model.flops_per_row / ms[len(ms) // 2] / 1e9:.0f} TFLOP/s)", flush=True)
