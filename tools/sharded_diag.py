"""Dev probe: where does the sharded guided frame differ from the unsharded one?"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mega_nerf_viewer_b200 as mnv

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tree = mnv.synth.make_tree(depth=6)
grid = mnv.synth.grid_for_world(world)
subs = [mnv.synth.make_mlp_weights(seed=11 + i) for i in range(world)]
gopt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8], use_guided_sampling=True, appearance_embedding=0)
w, h = 200, 113
cam = mnv.synth.default_camera(w, h, pose=3)
mn, mx = (-1, -1, -1), (1, 1, 1)
solo = mnv.multigpu.ReplicatedPipeline(tree, subs, grid, mn, mx)
want, rows_want = solo.guided_block(cam, gopt)
want = want.cpu().numpy()
sh = mnv.multigpu.ShardedGuided(tree, subs, grid, mn, mx, w, h, world=world)
got, rows = sh.guided_block(cam, gopt)
got = got.view(h, w, 4).cpu().numpy()
d = np.abs(got.astype(int) - want.astype(int)).max(-1)
print("rows", rows, rows_want, "bad>1:", int((d > 1).sum()), "max", d.max())
ys, xs = np.nonzero(d > 1)
# unsharded samples
g = solo.dt.guided_samples(cam, gopt, list(grid), list(mn), [2, 2, 2], capacity_rows=rows_want + 16)
off = g["offsets"].cpu().numpy(); z = g["z_vals"].cpu().numpy(); cl = g["cluster"].cpu().numpy(); rw = g["rows"].cpu().numpy()
# sharded per-cell samples
table = torch.empty((world, w * h, 4), device="cuda:0")
for c, dt in sh.trees.items():
    dt.guided_segment_probe(cam, sh._opt_for(gopt, c), out=table[c])
seg = {}
for c, dt in sh.trees.items():
    s = dt.guided_samples_segment(cam, sh._opt_for(gopt, c), list(grid), list(mn), [2, 2, 2], table, c, capacity_rows=rows_want + 16)
    seg[c] = (s["offsets"].cpu().numpy(), s["z_vals"].cpu().numpy(), s["cluster"].cpu().numpy(), s["rows"].cpu().numpy())
tb = table.cpu().numpy()
for y, x in list(zip(ys, xs))[:6]:
    i = y * w + x
    a, b = (0 if i == 0 else off[i - 1]), off[i]
    print(f"pixel ({x},{y}) d={d[y, x]} got={got[y, x]} want={want[y, x]}")
    print("  unsharded z", np.round(z[a:b], 5), "cluster", cl[a:b])
    for c in seg:
        o, zz, cc, rr = seg[c]
        a2, b2 = (0 if i == 0 else o[i - 1]), o[i]
        print(f"  cell {c} probe {tb[c, i]} z", np.round(zz[a2:b2], 5), "cluster", cc[a2:b2])
