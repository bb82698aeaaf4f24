"""Dev: time the native kernel only (no trackers / trackers) on the bench workload."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mega_nerf_viewer_b200 as mnv
depth = int(os.environ.get("DEPTH", "10")); W, H = 1920, 1080
tree = mnv.synth.make_tree(depth=depth); dt = mnv.DeviceTree(tree)
opt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8])
cams = [mnv.synth.default_camera(W, H, pose=i) for i in range(16)]
out = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
ts = torch.empty((W * H, 3), device="cuda"); tp = torch.empty((W * H, 3), device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(label, **kw):
    ms = []
    for i in range(40):
        flush.fill_(i & 255)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); dt.render(cams[i % 16], opt, out=out, **kw); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms = np.array(ms[8:]); print(f"{label}: mean {ms.mean():.3f} ms  min {ms.min():.3f}  max {ms.max():.3f}  -> {W*H/ms.mean()/1e3:.0f} Mrays/s", flush=True)
run("no-track"); run("track", to_split=ts, to_sample=tp)
