"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz.

The reference ships no tests, fixtures or golden images, so the
golden vectors are OUTPUTS OF THE REFERENCE ITSELF: its unmodified CUDA kernel
rebuilt for sm_100 (oracle/_ref/libref_render.so) plus the visit-log build
(libref_render_instr.so), run on a B200 through oracle/ref_driver.cpp.

Run on a GPU box (the reference kernel needs a GPU):

    gpurun -- 'python oracle/make_golden.py --out gpurun_out/golden'

then copy gpurun_out/golden/*.npz to tests/golden/ and commit.  Each file holds
the input tree (reference .npz schema), camera, RenderOptions fields and the
reference outputs: rgba, to_split, to_sample, per-ray visit hash / count / log.
While generating, the script also checks the CPU oracle and (if built) the
native CUDA path against the reference and prints one line per case.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402
import mega_nerf_viewer_b200 as mnv  # noqa: E402

W, H, LOG_CAP = 64, 36, 96


def cam_lookat(center, target, width=W, height=H, fx=None, up=(0.0, 0.0, 1.0)):
    center = np.asarray(center, np.float64)
    back = center - np.asarray(target, np.float64)
    back = (back / np.linalg.norm(back)).astype(np.float32)
    right = np.cross(np.asarray(up), back)
    right = (right / np.linalg.norm(right)).astype(np.float32)
    upv = np.cross(back, right).astype(np.float32)
    fx = fx if fx is not None else 1111.0 * width / 800.0
    return dict(width=width, height=height, fx=float(fx), fy=float(fx), cx=width / 2.0,
                cy=height / 2.0,
                c2w=np.concatenate([right, upv, back, center.astype(np.float32)]).astype(np.float32))


def cases():
    d = mnv.synth.default_camera
    yield "sh9_d5_pose0", dict(depth=5, fmt="SH9"), d(W, H, 0), {}
    yield "sh9_d5_pose5", dict(depth=5, fmt="SH9"), d(W, H, 5), {}
    yield "sh9_d6_inside", dict(depth=6, fmt="SH9"), cam_lookat((0.1, -0.2, 0.45), (0.6, 0.5, -0.3), fx=40.0), {}
    yield "sh9_d6_grazing", dict(depth=6, fmt="SH9"), cam_lookat((-2.5, 0.05, 0.02), (0.0, 0.0, 0.0), fx=120.0), {}
    yield "sh9_d5_far_miss", dict(depth=5, fmt="SH9"), cam_lookat((6.0, 5.0, 4.0), (0.0, 0.0, 0.0), fx=30.0), {}
    yield "rgba_d4", dict(depth=4, fmt="RGBA"), d(W, H, 2), {}
    yield "sh1_d4", dict(depth=4, fmt="SH1"), d(W, H, 3), {}
    yield "sh4_d4", dict(depth=4, fmt="SH4"), d(W, H, 7), {}
    yield "sh16_d4", dict(depth=4, fmt="SH16"), d(W, H, 9), {}
    yield "sh25_d4", dict(depth=4, fmt="SH25"), d(W, H, 11), {}
    yield "sh9_d5_sigma_hi", dict(depth=5, fmt="SH9"), d(W, H, 1), dict(sigma_thresh=30.0)
    yield "sh9_d5_sigma_zero", dict(depth=5, fmt="SH9"), d(W, H, 1), dict(sigma_thresh=-1.0)
    yield "sh9_d5_stop_half", dict(depth=5, fmt="SH9", sigma=(200.0, 900.0)), d(W, H, 4), dict(stop_thresh=0.5)
    yield "sh9_d5_dense_stop", dict(depth=5, fmt="SH9", sigma=(200.0, 900.0)), d(W, H, 6), {}
    yield "sh9_d5_step_big", dict(depth=5, fmt="SH9"), d(W, H, 8), dict(step_size=1e-2)
    yield "sh9_d5_bbox", dict(depth=5, fmt="SH9"), d(W, H, 10), dict(render_bbox=[0.2, 0.1, 0.3, 0.7, 0.8, 0.9])
    yield "sh9_d5_basis_1_4", dict(depth=5, fmt="SH9"), d(W, H, 12), dict(basis_minmax=[1, 4])
    yield "sh9_d5_rot", dict(depth=5, fmt="SH9"), d(W, H, 13), dict(rot_dirs=[0.3, -0.2, 0.9])
    yield "sh9_d5_depthmode", dict(depth=5, fmt="SH9"), d(W, H, 14), dict(render_depth=True)
    yield "sh9_d5_bg1", dict(depth=5, fmt="SH9"), d(W, H, 15), dict(background_brightness=1.0)
    yield "sh9_d5_maxdepth3", dict(depth=5, fmt="SH9"), d(W, H, 0), dict(max_depth=3)
    yield "sh9_d5_samplecap4", dict(depth=5, fmt="SH9"), d(W, H, 0), dict(max_sample_count=4)
    yield "sh9_d7_pose3", dict(depth=7, fmt="SH9"), d(W, H, 3), {}


def tree_sha(tree):
    import hashlib
    h = hashlib.sha256()
    for a in (tree.child, tree.parent, tree.data.view(np.uint16)):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def tree_fields(tree, key):
    """Trees are regenerated from (depth, format, sigma range) by the seeded
    generator and pinned by a sha256; only tiny trees are embedded verbatim."""
    f = dict(tree_spec_depth=np.int32(key[0]), tree_format=np.array(key[1]),
             tree_spec_sigma=np.array(key[2], np.float64), tree_sha256=np.array(tree_sha(tree)),
             tree_capacity=np.int64(tree.capacity))
    if tree.nbytes() < 200_000:
        f.update(tree_data=tree.data.view(np.uint16), tree_child=tree.child,
                 tree_parent=tree.parent, tree_depth=tree.depth, tree_scale=tree.scale,
                 tree_offset=tree.offset)
    return f


OPT_KEYS = ["step_size", "sigma_thresh", "stop_thresh", "background_brightness", "render_bbox",
            "basis_minmax", "rot_dirs", "render_depth", "max_depth", "max_sample_count"]


GRID, MINP, RNG = [2, 4], [-1.0, -1.0, -1.0], [2.0, 2.0, 2.0]


def nerf_cases():
    """Guided-sampling compositor (render_nerf_results_kernel, renderer_kernel.cu:294-327): the samples come from
    the reference's own get_samples_from_voxels, the per-sample values from a seeded generator."""
    d = mnv.synth.default_camera
    yield "nerf_sh9", dict(depth=5, fmt="SH9", sigma=(40.0, 300.0)), d(48, 27, 9), {}
    yield "nerf_rgba", dict(depth=5, fmt="RGBA", sigma=(40.0, 300.0)), d(48, 27, 3), {}
    yield "nerf_sh4", dict(depth=5, fmt="SH4", sigma=(40.0, 300.0)), d(48, 27, 6), {}
    yield "nerf_sh16", dict(depth=4, fmt="SH16", sigma=(40.0, 300.0)), d(48, 27, 12), {}
    yield "nerf_sh9_depthmode", dict(depth=5, fmt="SH9", sigma=(40.0, 300.0)), d(48, 27, 9), dict(render_depth=True)
    yield "nerf_sh9_rot_basis", dict(depth=5, fmt="SH9", sigma=(40.0, 300.0)), d(48, 27, 1), \
        dict(rot_dirs=[0.3, -0.2, 0.9], basis_minmax=[1, 6])


def nerf_values(V, D, seed=11):
    rng = np.random.default_rng(seed)
    values = rng.standard_normal((V, D + 1)).astype(np.float32)
    values[:, 3] = np.abs(values[:, 3]) * 30  # the column the reference reads sigma from (rt_core.cuh:365)
    return values


def make_nerf_goldens(out_dir):
    n_bad = 0
    for name, tspec, cam, okw in nerf_cases():
        key = (tspec["depth"], tspec["fmt"], tspec["sigma"])
        tree = mnv.synth.make_tree(depth=key[0], data_format=key[1], sigma_range=key[2])
        path = f"/tmp/golden_nerf_tree_{name}.npz"
        tree.save_npz(path)
        kw = dict(background_brightness=0.0, use_guided_sampling=True, max_guided_samples=24)
        kw.update(okw)
        if "basis_minmax" not in kw:
            kw["basis_minmax"] = [0, max(tree.basis_dim - 1, 0)]
        opt = O.default_options(**kw)
        ref = O.RefRenderer(path)
        g = ref.get_samples(cam, opt, GRID, MINP, RNG)
        off, z, _, _ = O.compact_samples(g["num_samples"], g["samples"], g["cluster"])
        assert off[-1] == z.shape[0] > 0
        values = nerf_values(z.shape[0], tree.data_dim)
        rimg = ref.render_nerf_results(cam, opt, values, z, off)
        assert rimg is not None, "reference compositor did not launch"
        out = dict(**tree_fields(tree, key),
                   cam_wh=np.array([cam["width"], cam["height"]], np.int32),
                   cam_intr=np.array([cam["fx"], cam["fy"], cam["cx"], cam["cy"]], np.float32),
                   cam_c2w=np.asarray(cam["c2w"], np.float32),
                   z_vals=z, offsets=off, values_seed=np.int32(11), values_sha=np.array(
                       __import__("hashlib").sha256(values.tobytes()).hexdigest()),
                   ref_rgba=rimg)
        for k in OPT_KEYS:
            v = getattr(opt, k)
            out["opt_" + k] = np.array(list(v) if hasattr(v, "__len__") else v)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **out)
        want = O.composite_nerf(tree, cam, opt, values, z, off)
        od = np.abs(want.astype(int) - rimg.astype(int))
        line = f"{name:22s} samples {z.shape[0]:6d} | oracle: maxabs {od.max()} ndiff {(od > 0).sum()}"
        try:
            import torch
            if torch.cuda.is_available() and os.path.exists(mnv.LIB_PATH):
                dt = mnv.DeviceTree(tree)
                img = dt.render_nerf_results(cam, mnv.default_options(**kw), torch.from_numpy(values).cuda(),
                                             torch.from_numpy(z).cuda(), torch.from_numpy(off).cuda()).cpu().numpy()
                md = np.abs(img.astype(int) - rimg.astype(int))
                n_bad += int(md.max() != 0)
                line += f" | native: maxabs {md.max()} ndiff {(md > 0).sum()} {'OK' if md.max() == 0 else 'MISMATCH'}"
                dt.close()
        except ImportError:
            pass
        print(line, flush=True)
        ref.close()
    print("native mismatching compositor cases:", n_bad)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "golden"))
    ap.add_argument("--only", default="all", choices=["all", "voxels", "nerf"])
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    if args.only in ("all", "nerf"):
        make_nerf_goldens(args.out)
    if args.only == "nerf":
        return
    try:
        import torch
        have_native = torch.cuda.is_available() and os.path.exists(mnv.LIB_PATH)
    except Exception:
        have_native = False
    trees = {}
    n_bad = 0
    for name, tspec, cam, okw in cases():
        key = (tspec["depth"], tspec["fmt"], tspec.get("sigma", (5.0, 50.0)))
        if key not in trees:
            t = mnv.synth.make_tree(depth=key[0], data_format=key[1], sigma_range=key[2])
            path = f"/tmp/golden_tree_{key[0]}_{key[1]}_{int(key[2][0])}.npz"
            t.save_npz(path)
            trees[key] = (t, path)
        tree, path = trees[key]
        kw = dict(background_brightness=0.0)
        kw.update(okw)
        if "basis_minmax" not in kw:  # VolumeRenderer::set, cuda_renderer.cpp:511-512
            kw["basis_minmax"] = [0, max(tree.basis_dim - 1, 0)]
        opt = O.default_options(**kw)
        ref = O.RefRenderer(path)
        refi = O.RefRenderer(path, instr=True)
        r = ref.render(cam, opt)
        ri = refi.render_logged(cam, opt, log_cap=LOG_CAP)
        assert np.array_equal(r["rgba"], ri["rgba"]), "instrumentation changed the image"
        out = dict(
            **tree_fields(tree, key),
            cam_wh=np.array([cam["width"], cam["height"]], np.int32),
            cam_intr=np.array([cam["fx"], cam["fy"], cam["cx"], cam["cy"]], np.float32),
            cam_c2w=np.asarray(cam["c2w"], np.float32),
            ref_rgba=r["rgba"], ref_to_split=r["to_split"], ref_to_sample=r["to_sample"],
            ref_hash=ri["hash"], ref_count=ri["count"], ref_log=ri["log"],
        )
        for k in OPT_KEYS:
            v = getattr(opt, k)
            out["opt_" + k] = np.array(list(v) if hasattr(v, "__len__") else v)
        np.savez_compressed(os.path.join(args.out, name + ".npz"), **out)

        # cross-checks
        o = O.render_voxels(tree, cam, opt, trackers=True, log_cap=LOG_CAP)
        od = np.abs(o["rgba"].astype(int) - r["rgba"].astype(int))
        oseq = ((o["hash"] != ri["hash"]) | (o["count"] != ri["count"])).sum()
        line = (f"{name:22s} rays {cam['width']*cam['height']} visits/ray {ri['count'].mean():6.1f} | oracle: "
                f"maxabs {od.max()} ndiff {(od>0).sum():4d} seq-diff {oseq:3d} "
                f"split {np.array_equal(o['to_split'], r['to_split'])} sample {np.array_equal(o['to_sample'], r['to_sample'])}")
        if have_native:
            import torch
            mopt = mnv.default_options(**kw)
            dt = mnv.DeviceTree(tree)
            m = dt.render_logged(cam, mopt, log_cap=LOG_CAP)
            P = cam["width"] * cam["height"]
            ts = torch.empty((P, 3), device="cuda")
            tp = torch.empty((P, 3), device="cuda")
            img = dt.render(cam, mopt, to_split=ts, to_sample=tp).cpu().numpy()
            md = np.abs(img.astype(int) - r["rgba"].astype(int))
            mseq = ((m["hash"] != ri["hash"]) | (m["count"] != ri["count"])).sum()
            ok = (md.max() == 0 and mseq == 0 and np.array_equal(ts.cpu().numpy(), r["to_split"])
                  and np.array_equal(tp.cpu().numpy(), r["to_sample"]) and np.array_equal(img, m["rgba"]))
            n_bad += 0 if ok else 1
            line += (f" | native: maxabs {md.max()} ndiff {(md>0).sum()} seq-diff {mseq} "
                     f"split {np.array_equal(ts.cpu().numpy(), r['to_split'])} "
                     f"sample {np.array_equal(tp.cpu().numpy(), r['to_sample'])} {'OK' if ok else 'MISMATCH'}")
            dt.close()
        print(line, flush=True)
        ref.close()
        refi.close()
    print("native mismatching cases:", n_bad)


if __name__ == "__main__":
    main()
