#!/usr/bin/env python
"""Dev: one 4K frame with candidate tracking, then the selection paths (for ncu launch lists / CUDA-event timing)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mega_nerf_viewer_b200 as mnv

import argparse
ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--height", type=int, default=2160)
a = ap.parse_args()
W, H = a.width, a.height
tree = mnv.synth.make_tree(depth=10)
dt = mnv.DeviceTree(tree)
opt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8])
P = W * H
ts = torch.empty((P, 3), device="cuda"); tp = torch.empty((P, 3), device="cuda")
dt.render(mnv.synth.default_camera(W, H, pose=3), opt, to_split=ts, to_sample=tp)
torch.cuda.synchronize()


def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, r


ms, (nodes, nc) = timed(lambda: mnv.select_candidates(ts, 4096, "split"))
print(f"select split: {ms:.3f} ms, candidates {nc}")
ms, (nodes, nc) = timed(lambda: mnv.select_candidates(tp, 4096, "sample"))
print(f"select sample: {ms:.3f} ms, candidates {nc}")
ms, rec = timed(lambda: mnv.vote_reduce(ts))
print(f"vote_reduce: {ms:.3f} ms, records {rec.shape[0]} of {P} rays")
ms, (nodes, nc) = timed(lambda: mnv.select_from_votes(rec, 4096, "split"))
print(f"select from records: {ms:.3f} ms, candidates {nc}")
