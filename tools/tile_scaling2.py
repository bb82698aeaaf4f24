"""Dev: where does the N-independent part of the frame time come from? (a) L2 flush, (b) longest ray."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mega_nerf_viewer_b200 as mnv
W, H = 1920, 1080
tree = mnv.synth.make_tree(depth=10); dt = mnv.DeviceTree(tree)
opt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8])
cams = [mnv.synth.default_camera(W, H, pose=i) for i in range(16)]
out = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(fn, n=40, do_flush=True, same_pose=False):
    ms = []
    for i in range(n):
        if do_flush: flush.fill_(i & 255)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(0 if same_pose else i); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return float(np.mean(ms[8:]))
for fl, sp in ((True, False), (False, False), (False, True)):
    full = run(lambda i: dt.render(cams[i % 16], opt, out=out), do_flush=fl, same_pose=sp)
    parts = [run(lambda i, m=m: dt.render_tiles(cams[i % 16], opt, out, 1920, 32, m, 0), 24, fl, sp) for m in (2, 4, 8, 16, 32)]
    print(f"flush={fl} same_pose={sp}: full {full:.3f}  1/2 {parts[0]:.3f}  1/4 {parts[1]:.3f}  1/8 {parts[2]:.3f}  1/16 {parts[3]:.3f}  1/32 {parts[4]:.3f}")
# the heaviest single CTA tile of pose 0, alone on the GPU
m = dt.render_logged(cams[0], opt)
c = m["count"].reshape(H, W)
tiles = c.reshape(H // 8, 8, W // 16, 16).max(axis=(1, 3))
ty, tx = np.unravel_index(tiles.argmax(), tiles.shape)
n_tiles = (H // 8) * (W // 16)
t_one = run(lambda i: dt.render_tiles(cams[0], opt, out, 16, 8, n_tiles, int(ty * (W // 16) + tx)), 24, False, True)
print(f"heaviest tile ({tiles.max()} visits max) alone, warm: {t_one:.3f} ms")
t_one = run(lambda i: dt.render_tiles(cams[0], opt, out, 16, 8, n_tiles, int(ty * (W // 16) + tx)), 24, True, True)
print(f"heaviest tile alone, L2 flushed: {t_one:.3f} ms")
t_e = run(lambda i: dt.render_tiles(cams[0], opt, out, 16, 8, n_tiles + 5, n_tiles + 1), 24, False, True)
print(f"empty launch (no tile selected): {t_e:.3f} ms")
