"""Dev: per-GPU cost of the image-tile partition on ONE GPU (what each of N GPUs would run)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mega_nerf_viewer_b200 as mnv
W, H = 1920, 1080
tree = mnv.synth.make_tree(depth=10); dt = mnv.DeviceTree(tree)
opt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8])
cams = [mnv.synth.default_camera(W, H, pose=i) for i in range(16)]
out = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(fn, n=40):
    ms = []
    for i in range(n):
        flush.fill_(i & 255)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(i); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return float(np.mean(ms[8:]))
full = run(lambda i: dt.render(cams[i % 16], opt, out=out))
print(f"full frame: {full:.3f} ms")
for mod in (2, 4, 8):
    for (tw, th) in ((1920, 8), (1920, 32), (64, 64), (16, 8), (128, 8)):
        t = [run(lambda i, r=r: dt.render_tiles(cams[i % 16], opt, out, tw, th, mod, r), 24) for r in range(min(mod, 4))]
        print(f"mod {mod} tile {tw}x{th}: per-rank ms {np.round(t, 3)}  max {max(t):.3f}  ideal {full / mod:.3f}  eff {full / mod / max(t):.2f}")
