"""Dev probe: single-process emulation of the multi-process sharded guided test (poses 0-3, 320x180)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mega_nerf_viewer_b200 as mnv
w, h = 320, 180
for world in (8, 4):
    tree = mnv.synth.make_tree(depth=6)
    grid = mnv.synth.grid_for_world(world)
    subs = [mnv.synth.make_mlp_weights(seed=11 + i) for i in range(world)]
    gopt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8], use_guided_sampling=True, appearance_embedding=0)
    mn, mx = (-1, -1, -1), (1, 1, 1)
    solo = mnv.multigpu.ReplicatedPipeline(tree, subs, grid, mn, mx)
    sh = mnv.multigpu.ShardedGuided(tree, subs, grid, mn, mx, w, h, world=world)
    for pose in range(4):
        cam = mnv.synth.default_camera(w, h, pose=pose)
        want, rw = solo.guided_block(cam, gopt)
        got, r = sh.guided_block(cam, gopt)
        d = np.abs(got.view(h, w, 4).cpu().numpy().astype(int) - want.cpu().numpy().astype(int))
        mse = float(np.mean(d.astype(np.float64) ** 2))
        print(world, pose, "rows", r, rw, "max", d.max(), "frac<=1", (d <= 1).mean(), "n>3", int((d.max(-1) > 3).sum()),
              "psnr", 99 if mse == 0 else 10 * np.log10(255 ** 2 / mse))
    sh.close(); solo.close()
