"""Dev (CPU only, oracle): how many lane-steps does the 8x4-tile-per-warp schedule waste, and what would lane-granular
refill from a warp-local / global queue recover?  Uses the per-ray leaf-visit counts of the CPU oracle."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mega_nerf_viewer_b200 as mnv
from oracle import oracle_py as O
W, H = 1920, 1080
tree = mnv.synth.make_tree(depth=10)
opt = O.default_options(background_brightness=0.0, basis_minmax=[0, 8])
for pose in (0, 5):
    r = O.render_voxels(tree, mnv.synth.default_camera(W, H, pose=pose), opt, stats=True)
    c = r["count"].reshape(H, W).astype(np.int64)
    # warp tiles 8 wide x 4 high
    t = c.reshape(H // 4, 4, W // 8, 8).transpose(0, 2, 1, 3).reshape(-1, 32)
    useful = t.sum()
    issued = (t.max(1) * 32).sum()
    print(f"pose {pose}: visits {useful/1e6:.1f} M, mean/ray {c.mean():.1f}; static 8x4 warps: lane-steps issued "
          f"{issued/1e6:.1f} M -> efficiency {useful/issued:.3f}")
    # lane refill simulation: persistent warp slots pull 8x4 blocks in tile order; a lane that finishes takes the next
    # pixel of the warp's current block queue once >= R lanes are idle (simplified: greedy list scheduling per warp
    # over K consecutive blocks = perfect refill -> steps = max(longest ray, ceil(sum/32)))
    for K in (2, 4, 8, 16):
        n = t.shape[0] // K * K
        g = t[:n].reshape(-1, K * 32)
        steps = np.maximum(g.max(1), -(-g.sum(1) // 32))
        print(f"   refill over {K} consecutive blocks (ideal): efficiency {g.sum()/(steps.sum()*32):.3f}")
