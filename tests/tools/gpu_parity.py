"""Dev script (GPU box): native kernel vs reference CUDA kernel vs CPU oracle.
Usage: python tools/gpu_parity.py [--depth 8] [--size 320x180] [--golden]"""
import argparse, os, sys, time, json
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import mega_nerf_viewer_b200 as mnv
from oracle import oracle_py as O

ap = argparse.ArgumentParser()
ap.add_argument("--depth", type=int, default=8)
ap.add_argument("--fmt", default="SH9")
ap.add_argument("--size", default="320x180")
ap.add_argument("--poses", type=int, default=4)
ap.add_argument("--big", default="")  # e.g. 1920x1080 timing run
ap.add_argument("--log-cap", type=int, default=64)
args = ap.parse_args()
W, H = map(int, args.size.split("x"))

print("device:", torch.cuda.get_device_name(0))
t0 = time.time()
tree = mnv.synth.make_tree(depth=args.depth, data_format=args.fmt)
print(f"tree depth {args.depth}: {tree.capacity} nodes, {tree.nbytes()/1e6:.1f} MB, gen {time.time()-t0:.1f}s")
npz = f"/tmp/tree_d{args.depth}_{args.fmt}.npz"
tree.save_npz(npz)

ref = O.RefRenderer(npz)
refi = O.RefRenderer(npz, instr=True) if O.ref_available(True) else None
dt = mnv.DeviceTree(tree)
# loader parity: the reference's cnpy/N3Tree::open arrays == generator arrays
rdata, rchild, rparent, rscale, roffset = ref.download()
print("ref loader == generator:", np.array_equal(rchild, tree.child), np.array_equal(rparent, tree.parent),
      np.array_equal(rdata.view(np.uint16), tree.data.view(np.uint16)), rscale, roffset)
mdata, mchild, mparent, msc = dt.download()
print("native round-trip:", np.array_equal(mchild, tree.child), np.array_equal(mparent, tree.parent),
      np.array_equal(mdata.view(np.uint16), tree.data.view(np.uint16)), int(msc.min()), int(msc.max()))

# point query
pts = np.random.default_rng(2).random((200000, 3)).astype(np.float32)
q = dt.query_points(pts).cpu().numpy()
print("query native == brute force:", np.array_equal(q, mnv.synth.brute_force_query(tree, pts)))

def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)

opt_o = O.default_options(background_brightness=0.0)
opt_m = mnv.default_options(background_brightness=0.0)
for pose in range(args.poses):
    cam = mnv.synth.default_camera(W, H, pose=pose, n_poses=args.poses)
    r = ref.render(cam, opt_o)
    mine = dt.render_logged(cam, opt_m, log_cap=args.log_cap)
    P = W * H
    ts = torch.full((P, 3), -7.0, device="cuda"); tp = torch.full((P, 3), -7.0, device="cuda")
    img = dt.render(cam, opt_m, to_split=ts, to_sample=tp); torch.cuda.synchronize()
    img = img.cpu().numpy()
    d = np.abs(r["rgba"].astype(int) - mine["rgba"].astype(int))
    print(f"pose {pose}: ref-vs-native(logged) maxabs {d.max()} n_diff {(d>0).sum()} psnr {psnr(r['rgba'], mine['rgba']):.1f} | "
          f"native logged==tracked img {np.array_equal(img, mine['rgba'])} | mean visits {mine['count'].mean():.1f} shaded {mine['shaded'].mean():.1f}")
    print("   trackers: split eq", np.array_equal(r["to_split"], ts.cpu().numpy()), " sample eq", np.array_equal(r["to_sample"], tp.cpu().numpy()))
    if not np.array_equal(r["to_split"], ts.cpu().numpy()):
        bad = np.nonzero((r["to_split"] != ts.cpu().numpy()).any(1))[0]
        print("   split mismatches", bad.size, "first", bad[:3], r["to_split"][bad[:3]], ts.cpu().numpy()[bad[:3]])
    if refi is not None:
        ri = refi.render_logged(cam, opt_o, log_cap=args.log_cap)
        print("   instr img == plain img:", np.array_equal(ri["rgba"], r["rgba"]))
        badh = np.nonzero((ri["hash"] != mine["hash"]) | (ri["count"] != mine["count"]))[0]
        print(f"   visit-seq mismatching rays: {badh.size} / {P}")
        for b in badh[:3]:
            print("    ray", b, "ref count", ri["count"][b], "mine", mine["count"][b])
            print("     ref ", ri["log"][b][:24])
            print("     mine", mine["log"][b][:24])
    # CPU oracle
    t1 = time.time(); o = O.render_voxels(tree, cam, opt_o, trackers=True, log_cap=args.log_cap); to = time.time() - t1
    d2 = np.abs(o["rgba"].astype(int) - r["rgba"].astype(int))
    if refi is not None:
        same = (o["hash"] == ri["hash"]) & (o["count"] == ri["count"])
        print(f"   oracle-vs-ref: maxabs {d2.max()} n_diff {(d2>0).sum()} psnr {psnr(o['rgba'], r['rgba']):.1f} visit-seq equal rays {same.mean()*100:.4f}% ({(~same).sum()} differ) cpu {to:.2f}s "
              f"split eq {np.array_equal(o['to_split'], r['to_split'])} sample eq {np.array_equal(o['to_sample'], r['to_sample'])}")
        for b in np.nonzero(~same)[0][:2]:
            print("    oracle ray", b, o["count"][b], ri["count"][b], o["log"][b][:16], ri["log"][b][:16])

if args.big:
    W2, H2 = map(int, args.big.split("x"))
    cam = mnv.synth.default_camera(W2, H2)
    r = ref.render(cam, opt_o, iters=12)
    print(f"REF {W2}x{H2}: ms/frame median {np.median(r['ms'][2:]):.3f}  all {np.round(r['ms'],3)}")
    out = torch.empty((H2, W2, 4), dtype=torch.uint8, device="cuda")
    ts = torch.empty((W2 * H2, 3), device="cuda"); tp = torch.empty((W2 * H2, 3), device="cuda")
    for label, kw in (("no-track", {}), ("track", dict(to_split=ts, to_sample=tp))):
        times = []
        for i in range(12):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); dt.render(cam, opt_m, out=out, **kw); e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        print(f"NATIVE {label} {W2}x{H2}: ms/frame median {np.median(times[2:]):.3f} all {np.round(times,3)}")
    img = out.cpu().numpy()
    d = np.abs(img.astype(int) - r["rgba"].astype(int))
    print(f"big frame parity: maxabs {d.max()} n_diff {(d>0).sum()} psnr {psnr(img, r['rgba']):.1f}")
    _, st = dt.render_frame_host(cam, opt_m, stats=True)
    print("stats", st, "alg bytes", st["visits"] * 6 + st["shaded_visits"] * 54 + st["rays"] * 4)
