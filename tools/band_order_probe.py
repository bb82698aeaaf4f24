"""Dev: does launching the 16x8-pixel tiles in row BANDS sorted by cost (heaviest band first, row-major inside a band)
shorten the tail of the traversal kernel?  Cost of a band = leaf visits of its rays, taken from (a) the same pose
(oracle: upper bound of what a predictor can give), (b) the previous pose of the 16-pose orbit (22.5 degrees away: what a
frame-to-frame predictor sees in the bench; a viewer moves far less between frames).
    python tools/band_order_probe.py [--band-rows 1|2|4] [--mill19]"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--band-rows", default="1,4")
ap.add_argument("--steps", type=int, default=48)
args = ap.parse_args()
import torch
import mega_nerf_viewer_b200 as mnv

W, H = 1920, 1080
tree = mnv.synth.make_tree(depth=10, data_format="SH9")
P = W * H
cams = [mnv.synth.default_camera(W, H, pose=i, n_poses=16) for i in range(16)]
opt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8])
dt = mnv.DeviceTree(tree, device=0)
out = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
ts = torch.empty((P, 3), device="cuda"); tp = torch.empty((P, 3), device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tx, ty = (W + 15) // 16, (H + 7) // 8
# per-tile cost of every pose (leaf visits + 8 x shaded visits, roughly instructions)
cost = []
for i in range(16):
    m = dt.render_logged(cams[i], opt, log_cap=0)
    c = (m["count"].astype(np.int64) + 8 * m["shaded"].astype(np.int64)).reshape(H, W)
    c = np.pad(c, ((0, ty * 8 - H), (0, tx * 16 - W)))
    cost.append(c.reshape(ty, 8, tx, 16).sum((1, 3)))


def order_from(tile_cost, band_rows):
    if tile_cost is None:
        return None
    nb = (ty + band_rows - 1) // band_rows
    band_cost = np.array([tile_cost[b * band_rows:(b + 1) * band_rows].sum() for b in range(nb)])
    bands = np.argsort(-band_cost, kind="stable")
    ids = np.concatenate([np.arange(b * band_rows * tx, min((b + 1) * band_rows, ty) * tx) for b in bands])
    return torch.from_numpy(ids.astype(np.int32)).cuda()


def timed(orders):
    ms = []
    for i in range(args.steps + 3):
        dt.set_tile_order(orders[i % 16])
        flush.fill_(i & 0xff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dt.render(cams[i % 16], opt, out=out, to_split=ts, to_sample=tp)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ms.append(e0.elapsed_time(e1))
    return round(float(np.mean(ms)), 4)


res = {"rows": timed([None] * 16)}
rev = torch.arange(tx * ty - 1, -1, -1, dtype=torch.int32).reshape(ty, tx).flip(1).reshape(-1).contiguous().cuda()  # bottom band first, row-major inside
res["bottom_up"] = timed([rev] * 16)
for br in [int(x) for x in args.band_rows.split(",")]:
    res[f"bands{br}_same_pose"] = timed([order_from(cost[i], br) for i in range(16)])
    res[f"bands{br}_prev_pose"] = timed([order_from(cost[(i - 1) % 16], br) for i in range(16)])
res["rows_again"] = timed([None] * 16)
print(json.dumps(res))
