"""Dev: frame time of the candidate-tracking and plain traversal variants over the bench orbit (L2 flushed)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mega_nerf_viewer_b200 as mnv
W, H = 1920, 1080
tree = mnv.synth.make_tree(depth=10); dt = mnv.DeviceTree(tree)
opt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8])
out = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = torch.empty((W * H, 3), device="cuda"); tp = torch.empty((W * H, 3), device="cuda")
cams = [mnv.synth.default_camera(W, H, pose=i, n_poses=16) for i in range(16)]
def run(n, **kw):
    ms = []
    for i in range(n):
        flush.fill_(i & 255)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); dt.render(cams[i % 16], opt, out=out, **kw); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return float(np.mean(ms[16:]))
a = dt.render(cams[3], opt).clone(); b = dt.render(cams[3], opt, to_split=ts, to_sample=tp)
print("images equal:", bool(torch.equal(a, b)))
print("track %.4f ms   plain %.4f ms" % (run(64, to_split=ts, to_sample=tp), run(64)))
