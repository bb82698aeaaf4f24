"""Dev: how much does launching the heaviest CTA tiles first buy? (oracle order from the true visit counts)"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mega_nerf_viewer_b200 as mnv
W, H = 1920, 1080
tree = mnv.synth.make_tree(depth=10); dt = mnv.DeviceTree(tree)
opt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8])
out = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = torch.empty((W * H, 3), device="cuda"); tp = torch.empty((W * H, 3), device="cuda")
def run(cam, n=24, **kw):
    ms = []
    for i in range(n):
        flush.fill_(i & 255)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); dt.render(cam, opt, out=out, **kw); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return float(np.mean(ms[4:]))
for pose in (0, 5, 11):
    cam = mnv.synth.default_camera(W, H, pose=pose)
    m = dt.render_logged(cam, opt)
    c = m["count"].reshape(H // 8, 8, W // 16, 16)
    res = {}
    dt.set_tile_order(None); ref_img = dt.render(cam, opt).clone(); res["row-major"] = (run(cam), run(cam, to_split=ts, to_sample=tp))
    for name, key in (("max-visits first", c.max(axis=(1, 3)).ravel()), ("sum-visits first", c.sum(axis=(1, 3)).ravel()),
                      ("max, coarse 8 buckets", (c.max(axis=(1, 3)).ravel() // 64))):
        order = torch.from_numpy(np.argsort(-key.astype(np.int64), kind="stable").astype(np.int32)).cuda()
        dt.set_tile_order(order)
        assert torch.equal(dt.render(cam, opt), ref_img)
        res[name] = (run(cam), run(cam, to_split=ts, to_sample=tp))
    rnd = torch.from_numpy(np.random.default_rng(0).permutation(c.shape[0] * c.shape[2]).astype(np.int32)).cuda()
    dt.set_tile_order(rnd); res["random"] = (run(cam), run(cam, to_split=ts, to_sample=tp))
    print(f"pose {pose}: " + "  ".join(f"{k}: {a:.3f}/{b:.3f}" for k, (a, b) in res.items()), flush=True)
