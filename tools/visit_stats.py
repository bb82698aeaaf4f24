"""Dev: distribution of leaf visits per ray on the bench workload (where is the kernel's tail?)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mega_nerf_viewer_b200 as mnv
W, H = 1920, 1080
tree = mnv.synth.make_tree(depth=10); dt = mnv.DeviceTree(tree)
opt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8])
for pose in (0, 5):
    m = dt.render_logged(mnv.synth.default_camera(W, H, pose=pose), opt)
    c = m["count"].reshape(H, W)
    print(f"pose {pose}: mean {c.mean():.1f} p50 {np.percentile(c,50):.0f} p90 {np.percentile(c,90):.0f} p99 {np.percentile(c,99):.0f} p99.9 {np.percentile(c,99.9):.0f} max {c.max()}")
    rows = c.reshape(H // 8, 8, W).max(axis=(1, 2))
    print("  max visits per 8-row band (every 10th band):", rows[::10].tolist())
    tiles = c.reshape(H // 8, 8, W // 16, 16).sum(axis=(1, 3))
    print("  CTA-tile visit sums: mean %.0f max %d; per band-of-tiles mean (every 10th):" % (tiles.mean(), tiles.max()), tiles.mean(1)[::10].astype(int).tolist())
