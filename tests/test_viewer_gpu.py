"""GPU tests of viewer::VolumeRenderer (csrc/viewer/renderer.cpp) through the headless driver:
the C++ frame orchestration of Impl::render (cuda_renderer.cpp:68-163) must produce exactly what the
same kernels produce when driven call by call from Python, and the Camera must match the
reference's own camera.cpp (run through oracle/_ref on this box)."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def run(mnv, *args):
    r = subprocess.run([mnv.HEADLESS_BIN, *map(str, args)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads([l for l in r.stdout.strip().splitlines() if l.startswith("{")][-1])


def cam_from(j, w, h):
    return dict(width=w, height=h, fx=j["fx"], fy=j["fx"], cx=j["cx"], cy=j["cy"], c2w=np.asarray(j["c2w"], np.float32))


def test_headless_octree_frame_is_bit_exact(mnv, tmp_path):
    tree = mnv.synth.make_tree(depth=7)
    path = tmp_path / "t.npz"
    tree.save_npz(str(path), compressed=True)
    raw = tmp_path / "f.rgba"
    w, h = 640, 360
    j = run(mnv, path, "--width", w, "--height", h, "--frames", 5, "--raw", raw, "--out", tmp_path / "f.ppm")
    assert j["capacity"] == tree.capacity and j["frames"] == 5
    got = np.fromfile(raw, np.uint8).reshape(h, w, 4)
    assert j["frame_hash"] == mnv.bytes_checksum(got)
    dt = mnv.DeviceTree(tree)
    # VolumeRenderer::set narrows basis_minmax to the tree's basis (cuda_renderer.cpp:511-512)
    opt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8])
    want = dt.render(cam_from(j, w, h), opt).cpu().numpy()
    assert np.array_equal(got, want)
    assert (got[..., :3].max() > 0) and (got[..., 3] == 255).all()
    ppm = (tmp_path / "f.ppm").read_bytes()
    assert ppm.startswith(b"P6\n640 360\n255\n") and len(ppm) == 15 + w * h * 3
    dt.close()


@pytest.mark.parametrize("bg", [0.0, 1.0])
def test_headless_interop_surfaces_equal_the_offscreen_frame(mnv, tmp_path, bg):
    """VolumeRenderer::set_interop_surfaces: frames composited in place over cleared RGBA8 / R32F cudaArray surfaces
    (the GL viewer's presentation path, offscreen = false) equal the headless offscreen frames."""
    tree = mnv.synth.make_tree(depth=6)
    path = tmp_path / "t.npz"
    tree.save_npz(str(path))
    a, b = tmp_path / "a.rgba", tmp_path / "b.rgba"
    ja = run(mnv, path, "--width", 320, "--height", 180, "--frames", 3, "--bg", bg, "--raw", a)
    jb = run(mnv, path, "--width", 320, "--height", 180, "--frames", 3, "--bg", bg, "--raw", b, "--interop")
    fa, fb = np.fromfile(a, np.uint8), np.fromfile(b, np.uint8)
    assert np.array_equal(fa, fb) and ja["frame_hash"] == jb["frame_hash"]
    assert fa.reshape(180, 320, 4)[..., :3].std() > 0


def make_model(mnv, path, n_sub=1, grid=(1, 1)):
    import torch
    from mlp_reference import MegaNerfMLP

    subs = []
    for s in range(n_sub):
        torch.manual_seed(3 + s)
        subs.append(MegaNerfMLP().export())
    mnv.save_model_container(str(path), subs, grid_dim=grid, min_position=(-1, -1, -1), max_position=(1, 1, 1))
    return subs


def test_headless_refinement_grows_the_tree(mnv, tmp_path):
    tree = mnv.synth.make_tree(depth=6)
    path, mpath = tmp_path / "t.npz", tmp_path / "m.npz"
    tree.save_npz(str(path))
    make_model(mnv, mpath, n_sub=2, grid=(1, 2))
    j = run(mnv, path, "--model", mpath, "--width", 480, "--height", 270, "--frames", 6, "--poses", 1,
            "--use_splitting", "--max_tree_capacity", tree.capacity + 40000)
    assert j["nodes_added"] > 0
    assert j["capacity"] > tree.capacity
    # same seed -> same refinement -> same frame
    j2 = run(mnv, path, "--model", mpath, "--width", 480, "--height", 270, "--frames", 6, "--poses", 1,
             "--use_splitting", "--max_tree_capacity", tree.capacity + 40000)
    assert j2["frame_hash"] == j["frame_hash"] and j2["capacity"] == j["capacity"]


def test_headless_saves_the_refined_tree(mnv, tmp_path):
    """--save after refinement: the file holds the grown tree and renders like the live one."""
    tree = mnv.synth.make_tree(depth=6)
    path, mpath, out = tmp_path / "t.npz", tmp_path / "m.npz", tmp_path / "refined.npz"
    tree.save_npz(str(path))
    make_model(mnv, mpath)
    raw = tmp_path / "f.rgba"
    j = run(mnv, path, "--model", mpath, "--width", 320, "--height", 180, "--frames", 3, "--poses", 1,
            "--use_splitting", "--max_tree_capacity", tree.capacity + 20000, "--save", out)
    refined = mnv.HostTree.load_npz(str(out))
    assert refined.capacity == j["capacity"] > tree.capacity
    assert np.array_equal(refined.child[: tree.capacity] != 0, tree.child != 0) is False  # some leaves were split
    # a fresh, non-refining run on the saved tree reproduces a frame of the refined scene
    j2 = run(mnv, out, "--width", 320, "--height", 180, "--frames", 1, "--poses", 1, "--raw", raw)
    assert j2["capacity"] == refined.capacity
    dt = mnv.DeviceTree(refined)
    opt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8])
    want = dt.render(cam_from(j2, 320, 180), opt).cpu().numpy()
    assert np.array_equal(np.fromfile(raw, np.uint8).reshape(180, 320, 4), want)
    dt.close()


@pytest.mark.parametrize("n_retain", [0, 2])
def test_vq_tree_decoded_on_the_gpu(mnv, tmp_path, n_retain):
    """quant_colors / quant_map / data_retained / sigma (n3tree.cpp:109-175) go to the device
    compressed and are decoded there — through the C-ABI (mnv_tree_create_vq) and through viewer::N3Tree::open +
    move_to_device — and equal the numpy decode bit for bit."""
    tree = mnv.synth.make_tree(depth=5)
    cap, nb = tree.capacity, 9
    n_q = nb - n_retain
    rng = np.random.default_rng(21 + n_retain)
    book = rng.standard_normal((n_q, 65536, 3)).astype(np.float16)
    qmap = rng.integers(0, 65536, (n_q, cap, 8), dtype=np.uint16)
    retained = rng.standard_normal((n_retain, cap, 8, 3)).astype(np.float16)
    sigma = np.ascontiguousarray(tree.data[..., -1])
    want = np.zeros((cap, 8, 28), np.float16)
    for b in range(n_q):
        col = book[b][qmap[b]]
        for ch in range(3):
            want[:, :, ch * nb + n_retain + b] = col[..., ch]
    for b in range(n_retain):
        for ch in range(3):
            want[:, :, ch * nb + b] = retained[b][..., ch]
    want[:, :, 27] = sigma
    # C-ABI
    dt = mnv.DeviceTree(tree, vq=dict(quant_colors=book, quant_map=qmap, data_retained=retained, sigma=sigma))
    data, child, parent, counts = dt.download()
    assert np.array_equal(data.view(np.uint16), want.view(np.uint16))
    assert np.array_equal(child, tree.child) and (counts == 8).all()
    dt.close()
    # C++ API: the file stays compressed on the host, the frame equals the one of the decoded tree
    path, raw = tmp_path / "vq.npz", tmp_path / "f.rgba"
    np.savez(path, data_dim=np.int64(28), data_format=np.array("SH9"), invradius3=tree.scale, offset=tree.offset,
             child=tree.child.reshape(cap, 2, 2, 2), parent_depth=np.stack([tree.parent, tree.depth], 1).astype(np.int32),
             quant_colors=book, quant_map=qmap.reshape(n_q, cap, 2, 2, 2), sigma=sigma.reshape(cap, 2, 2, 2),
             **({"data_retained": retained.reshape(n_retain, cap, 2, 2, 2, 3)} if n_retain else {}))
    j = run(mnv, path, "--width", 320, "--height", 180, "--frames", 1, "--poses", 1, "--raw", raw)
    decoded = mnv.HostTree(N=2, data_dim=28, data_format="SH9", child=tree.child, parent=tree.parent, depth=tree.depth,
                           data=want, scale=tree.scale, offset=tree.offset)
    dt = mnv.DeviceTree(decoded)
    opt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8])
    img = dt.render(cam_from(j, 320, 180), opt).cpu().numpy()
    assert np.array_equal(np.fromfile(raw, np.uint8).reshape(180, 320, 4), img)
    assert img[..., :3].max() > 0
    dt.close()


def test_headless_replica_group_frames_and_refinement(mnv, tmp_path):
    """viewer::VolumeRenderer::set_devices (C++ multi-GPU host path, csrc/mnv_group.cu): the gathered frame equals the
    one-GPU frame bit for bit — headless and through the interop surface — and refinement across three replicas
    reproduces itself (same seed, same frame hash and capacity)."""
    tree = mnv.synth.make_tree(depth=6)
    path, mpath = tmp_path / "t.npz", tmp_path / "m.npz"
    tree.save_npz(str(path))
    make_model(mnv, mpath)
    one = run(mnv, path, "--width", 400, "--height", 232, "--frames", 2, "--poses", 1)
    grp = run(mnv, path, "--width", 400, "--height", 232, "--frames", 2, "--poses", 1, "--replicas", 3)
    assert grp["frame_hash"] == one["frame_hash"]
    grp_i = run(mnv, path, "--width", 400, "--height", 232, "--frames", 2, "--poses", 1, "--replicas", 3, "--interop")
    one_i = run(mnv, path, "--width", 400, "--height", 232, "--frames", 2, "--poses", 1, "--interop")
    assert grp_i["frame_hash"] == one_i["frame_hash"] == one["frame_hash"]
    args = ("--model", mpath, "--width", 400, "--height", 232, "--frames", 4, "--poses", 1, "--use_splitting",
            "--max_tree_capacity", tree.capacity + 40000, "--replicas", 3)
    a, b = run(mnv, path, *args), run(mnv, path, *args)
    # (the two warm-up frames refine too: capacity grows by more than the timed frames' count)
    assert a["nodes_added"] > 0 and tree.capacity + a["nodes_added"] <= a["capacity"] <= tree.capacity + 40000
    assert a["frame_hash"] == b["frame_hash"] and a["capacity"] == b["capacity"]


def test_headless_prunes_when_full(mnv, tmp_path):
    """max_tree_capacity - capacity < split_batch_size triggers Impl::prune_tree
    (cuda_renderer.cpp:146-151).  Visit tracking only starts once capacity > 3/4 max or after a
    prune (:99-100), so — exactly like the reference — the first prune pass of a tree that filled up
    quickly sees an empty tracker and drops everything but the root; the renderer must survive it."""
    tree = mnv.synth.make_tree(depth=6)
    path, mpath = tmp_path / "t.npz", tmp_path / "m.npz"
    tree.save_npz(str(path))
    make_model(mnv, mpath)
    j = run(mnv, path, "--model", mpath, "--width", 480, "--height", 270, "--frames", 8, "--poses", 1,
            "--use_splitting", "--max_tree_capacity", tree.capacity + 6000)
    assert j["capacity"] <= tree.capacity + 6000
    assert j["capacity"] >= 1


def test_headless_guided_sampling_matches_python_pipeline(mnv, tmp_path):
    import torch

    tree = mnv.synth.make_tree(depth=6)
    path, mpath, raw = tmp_path / "t.npz", tmp_path / "m.npz", tmp_path / "g.rgba"
    tree.save_npz(str(path))
    subs = make_model(mnv, mpath, n_sub=2, grid=(1, 2))
    w, h = 320, 180
    j = run(mnv, path, "--model", mpath, "--width", w, "--height", h, "--frames", 3, "--poses", 1,
            "--use_guided_sampling", "--raw", raw)
    got = np.fromfile(raw, np.uint8).reshape(h, w, 4)
    # frames 2.. reuse the results of the first (camera static): rows are counted once per camera move
    assert j["guided_rows"] == 0
    dt = mnv.DeviceTree(tree)
    model = mnv.MlpModel(subs, grid_dim=(1, 2), min_position=(-1, -1, -1), max_position=(1, 1, 1))
    opt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8], use_guided_sampling=True,
                              appearance_embedding=0)
    cam = cam_from(j, w, h)
    g = dt.guided_samples(cam, opt, (1, 2), (-1, -1, -1), (2, 2, 2), capacity_rows=w * h * 64)
    assert g["total"] > 0
    vals = torch.zeros((g["total"], tree.data_dim + 1), device="cuda")
    model.query_submodules(g["cluster"], g["rows"], vals)
    want = dt.render_nerf_results(cam, opt, vals, g["z_vals"], g["offsets"], sigma_col=tree.data_dim - 1)
    assert np.array_equal(got, want.cpu().numpy())
    model.close()
    dt.close()


def test_camera_matches_reference_camera_cpp(mnv):
    """The reference's own src/camera.cpp (inside oracle/_ref/libref_render.so) vs viewer::Camera,
    same script; also (re)writes tests/golden/camera_trace.json for the CPU suite."""
    so = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_render.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built")
    import torch  # noqa: F401  (libref links libtorch)

    L = C.CDLL(so)
    w, h = 800, 600
    out = np.zeros((3, 48), np.float32)
    L.ref_camera_trace.argtypes = [C.c_int, C.c_int, C.c_float, C.c_void_p]
    assert L.ref_camera_trace(w, h, C.c_float(1111.0 * (w / 800.0)), out.ctypes.data) == 0
    r = subprocess.run([mnv.HEADLESS_BIN, "--selftest-camera", "--width", str(w), "--height", str(h)],
                       capture_output=True, text=True)
    steps = [json.loads(l) for l in r.stdout.strip().splitlines()]
    for s, ref in zip(steps, out):
        got = np.asarray(s["transform"] + s["K"] + s["w2c"] + [s["fx"], s["fy"], s["cx"], s["cy"]], np.float32)
        assert np.allclose(got, ref, rtol=2e-6, atol=2e-6), np.abs(got - ref).max()
    gdir = os.environ.get("MNV_GOLDEN_OUT")
    if gdir:
        os.makedirs(gdir, exist_ok=True)
        with open(os.path.join(gdir, "camera_trace.json"), "w") as f:
            json.dump(dict(width=w, height=h, steps=[[float(v) for v in row] for row in out]), f)
