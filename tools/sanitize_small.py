"""Dev: one small pass over every kernel family, meant to run under `compute-sanitizer --tool memcheck`
(tools/sessions/gpu_sanitize.sh).  Sizes are tiny: the sanitizer slows kernels down 10-100x."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mega_nerf_viewer_b200 as mnv

what = set(sys.argv[1:]) or {"render", "guided", "refine", "split", "mlp"}
GRID, MINP, MAXP, RNG = [1, 2], [-1.0] * 3, [1.0] * 3, [2.0] * 3
tree = mnv.synth.make_tree(depth=5)
w, h = 96, 56
P = w * h
cam = mnv.synth.default_camera(w, h, pose=2)
opt = mnv.default_options(background_brightness=0.5, basis_minmax=[0, 8])
dt = mnv.DeviceTree(tree, max_capacity=tree.capacity + 512)

if "render" in what:
    ts, tp = torch.empty((P, 3), device="cuda"), torch.empty((P, 3), device="cuda")
    dt.render(cam, opt, to_split=ts, to_sample=tp)
    dt.render(cam, opt)
    pts = torch.rand((5000, 3), device="cuda")
    if hasattr(dt, "query_points"):
        dt.query_points(pts)
    nodes, _ = mnv.select_candidates(ts, 64, "split")
    mnv.select_candidates(tp, 64, "sample")
    torch.cuda.synchronize()
    print("render ok", nodes.shape[0])

subs = [mnv.synth.make_mlp_weights(seed=3 + i) for i in range(2)]
if "mlp" in what or "guided" in what or "refine" in what:
    model = mnv.MlpModel(subs, grid_dim=GRID, min_position=MINP, max_position=MAXP)

if "mlp" in what:
    x = torch.rand((1000, model.in_dim), device="cuda")
    x[:, -1] = 0
    y = model.forward(x, 1)
    torch.cuda.synchronize()
    print("mlp ok", float(y.abs().mean()))

if "guided" in what:
    gopt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8], use_guided_sampling=True,
                               appearance_embedding=0)
    g = dt.guided_samples(cam, gopt, GRID, MINP, RNG, capacity_rows=P * 24)
    vals = torch.empty((max(g["total"], 1), tree.data_dim + 1), device="cuda")
    model.query_submodules(g["cluster"], g["rows"], vals)
    dt.render_nerf_results(cam, gopt, vals, g["z_vals"], g["offsets"], sigma_col=tree.data_dim - 1)
    torch.cuda.synchronize()
    print("guided ok", g["total"])

if "refine" in what:
    ropt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8], use_splitting=True, appearance_embedding=0,
                               split_batch_size=32)
    ts, tp = torch.empty((P, 3), device="cuda"), torch.empty((P, 3), device="cuda")
    dt.render(cam, ropt, to_split=ts, to_sample=tp)
    nodes, _ = mnv.select_candidates(ts, 32, "split")
    k, c = nodes.shape[0], ropt.samples_per_corner
    samples = torch.rand((k * 8, c, 4), device="cuda")
    cluster = torch.zeros((k * 8, c), dtype=torch.int16, device="cuda")
    dt.add_children(ropt, nodes, samples, cluster, GRID, MINP, RNG)
    res = torch.empty((k * 8 * c, tree.data_dim + 1), device="cuda")
    model.query_submodules(cluster.view(-1), samples.view(-1, 4), res)
    dt.commit_children(ropt, k, res.view(k * 8, c, -1))
    dt.render(cam, ropt)
    torch.cuda.synchronize()
    print("refine ok", k, dt.capacity)

if "split" in what:
    sp = mnv.multigpu.SubmoduleSplit(tree, w, h, world=2)
    sp.render_full_single(cam, opt)
    sp.close()
    gopt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8], use_guided_sampling=True,
                               appearance_embedding=0)
    sh = mnv.multigpu.ShardedGuided(tree, subs, (1, 2), MINP, MAXP, w, h, world=2)
    _, rows = sh.guided_block(cam, gopt)
    sh.close()
    torch.cuda.synchronize()
    print("split ok", rows)
print("done")
