#!/usr/bin/env python
"""Dev: A/B timing of traversal-kernel variants on the bench workload (1080p, depth-10 SH9 tree, 16-pose orbit).

    python tools/variant_bench.py [--lib path/to/libmnv_b200_X.so] [--anchor 0,5,6,7,8] [--mill19]

For every anchor level (MNV_ANCHOR_LEVEL is read when a tree is created; 0 = the round-1 path-cache march) it
prints one JSON line: mean ms per frame with and without candidate tracking (CUDA events, L2 flushed between
frames), a CRC of all 16 frames + both trackers (equal CRCs = bit-identical output), and the visit-hash CRC of the
logged variant.  The library under test comes from MNV_B200_LIB / --lib (compile-time variants are built by
tools/build_variants.sh)."""
import argparse, json, os, sys, zlib
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=None)
ap.add_argument("--anchor", default="0,6,7,8")
ap.add_argument("--steps", type=int, default=48)
ap.add_argument("--mill19", action="store_true")
ap.add_argument("--max-nodes", type=int, default=16_000_000)
ap.add_argument("--tag", default="")
ap.add_argument("--order", default="rows", choices=["rows", "morton", "cols", "snake"], help="launch order of the 16x8 CTA tiles")
args = ap.parse_args()
if args.lib:
    os.environ["MNV_B200_LIB"] = os.path.abspath(args.lib)
import torch
import mega_nerf_viewer_b200 as mnv

if args.mill19:
    W, H = 3840, 2160
    tree = mnv.synth.make_tree(depth=12, data_format="SH9", blocks_yz=(2, 4), block_depths=[12, 11, 11, 12, 11, 12, 12, 11],
                               max_nodes=args.max_nodes)
else:
    W, H = 1920, 1080
    tree = mnv.synth.make_tree(depth=10, data_format="SH9")
P = W * H
cams = [mnv.synth.default_camera(W, H, pose=i, n_poses=16) for i in range(16)]
opt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8])
dev = torch.device("cuda", 0)
out = torch.zeros((H, W, 4), dtype=torch.uint8, device=dev)
ts = torch.empty((P, 3), device=dev)
tp = torch.empty((P, 3), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(dt, trackers, steps):
    ms = []
    for i in range(steps + 3):
        flush.fill_(i & 0xff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dt.render(cams[i % 16], opt, out=out, to_split=ts if trackers else None, to_sample=tp if trackers else None)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ms.append(e0.elapsed_time(e1))
    return float(np.mean(ms)), float(np.min(ms))


for a in [int(x) for x in args.anchor.split(",")]:
    os.environ["MNV_ANCHOR_LEVEL"] = str(a)
    dt = mnv.DeviceTree(tree, device=0)
    if args.order != "rows":
        tx, ty = (W + 15) // 16, (H + 7) // 8
        yy, xx = np.mgrid[0:ty, 0:tx]
        ids = (yy * tx + xx).ravel()
        if args.order == "cols":
            key = (xx * ty + yy).ravel()
        elif args.order == "snake":  # 8x8-tile super blocks, row-major inside
            key = (((yy // 8) * ((tx + 7) // 8) + xx // 8) * 64 + (yy % 8) * 8 + xx % 8).ravel()
        else:
            def spread(v):
                v = v.astype(np.int64); r = np.zeros_like(v)
                for b in range(10):
                    r |= ((v >> b) & 1) << (2 * b)
                return r
            key = (spread(xx) | (spread(yy) << 1)).ravel()
        dt.set_tile_order(torch.from_numpy(ids[np.argsort(key, kind="stable")].astype(np.int32)).cuda())
    crc = 0
    for i in range(16 if not args.mill19 else 4):
        dt.render(cams[i], opt, out=out, to_split=ts, to_sample=tp)
        torch.cuda.synchronize()
        for t in (out, ts, tp):
            crc = zlib.crc32(t.cpu().numpy().tobytes(), crc)
    vh = 0
    if not args.mill19:
        m = dt.render_logged(cams[3], opt, log_cap=0)
        vh = zlib.crc32(np.ascontiguousarray(m["hash"]).tobytes(), zlib.crc32(np.ascontiguousarray(m["count"]).tobytes()))
    t_tr, t_tr_min = timed(dt, True, args.steps)
    t_nt, t_nt_min = timed(dt, False, args.steps)
    print(json.dumps({"tag": args.tag, "lib": os.path.basename(mnv.LIB_PATH), "anchor_level": a, "ms_track": round(t_tr, 4),
                      "ms_track_min": round(t_tr_min, 4), "ms_notrack": round(t_nt, 4), "ms_notrack_min": round(t_nt_min, 4),
                      "mrays_track": round(P / t_tr / 1e3, 1), "frames_crc": f"{crc:08x}", "visit_crc": f"{vh:08x}",
                      "res": [W, H], "nodes": int(tree.capacity)}), flush=True)
    dt.close()
