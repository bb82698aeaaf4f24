"""Dev: tile order predicted from the PREVIOUS pose of the orbit (what a frame-to-frame scheduler would know)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mega_nerf_viewer_b200 as mnv
W, H = 1920, 1080
tree = mnv.synth.make_tree(depth=10); dt = mnv.DeviceTree(tree)
opt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8])
out = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = torch.empty((W * H, 3), device="cuda"); tp = torch.empty((W * H, 3), device="cuda")
def run(cam, n=20, **kw):
    ms = []
    for i in range(n):
        flush.fill_(i & 255)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); dt.render(cam, opt, out=out, **kw); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return float(np.mean(ms[4:]))
def key_of(pose):
    m = dt.render_logged(mnv.synth.default_camera(W, H, pose=pose), opt)
    c = m["count"].reshape(H // 8, 8, W // 16, 16)
    return c.max(axis=(1, 3)).ravel().astype(np.int64), c.sum(axis=(1, 3)).ravel().astype(np.int64)
def order(key):
    return torch.from_numpy(np.argsort(-key, kind="stable").astype(np.int32)).cuda()
for pose in (1, 6, 12):
    cam = mnv.synth.default_camera(W, H, pose=pose)
    kmax, ksum = key_of(pose)
    pmax, psum = key_of(pose - 1)
    res = {}
    dt.set_tile_order(None); res["row-major"] = (run(cam), run(cam, to_split=ts, to_sample=tp))
    for name, k in (("oracle max", kmax), ("prev-pose max", pmax), ("prev-pose sum", psum), ("prev max /32", pmax // 32),
                    ("prev max>=p90 first", (pmax >= np.percentile(pmax, 90)).astype(np.int64))):
        dt.set_tile_order(order(k)); res[name] = (run(cam), run(cam, to_split=ts, to_sample=tp))
    print(f"pose {pose}: " + "  ".join(f"{k}: {a:.3f}/{b:.3f}" for k, (a, b) in res.items()), flush=True)
