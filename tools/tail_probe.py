"""Dev: the heaviest CTA tile of pose 0 alone on the GPU (the kernel's latency floor), for ncu."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mega_nerf_viewer_b200 as mnv
W, H = 1920, 1080
tree = mnv.synth.make_tree(depth=10); dt = mnv.DeviceTree(tree)
opt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8])
cam = mnv.synth.default_camera(W, H, pose=0)
out = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
m = dt.render_logged(cam, opt)
c = m["count"].reshape(H, W)
tiles = c.reshape(H // 8, 8, W // 16, 16).max(axis=(1, 3))
ty, tx = np.unravel_index(tiles.argmax(), tiles.shape)
n_tiles = (H // 8) * (W // 16)
tile_counts = c.reshape(H // 8, 8, W // 16, 16)[ty, :, tx, :]
print("tile", ty, tx, "visits per ray: max", tile_counts.max(), "mean", tile_counts.mean())
for i in range(4):
    dt.render_tiles(cam, opt, out, 16, 8, n_tiles, int(ty * (W // 16) + tx))
torch.cuda.synchronize()
