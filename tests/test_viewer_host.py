"""CPU tests of the host-side mirror of the reference's C++ API (csrc/viewer/): the .npz loader
behind viewer::N3Tree::open (reference: src/n3tree/n3tree.cpp:16-205 + the vendored cnpy),
pack/unpack_index, gen_wireframe (:248-336) and viewer::Camera (src/camera.cpp), exercised through
the `mnv_headless --selftest-*` modes of the headless driver.  No GPU needed."""
import json
import os
import subprocess
import zipfile

import numpy as np
import pytest


def run(mnv, *args):
    return subprocess.run([mnv.HEADLESS_BIN, *map(str, args)], capture_output=True, text=True, timeout=120)


def last_json(out: str):
    return json.loads([l for l in out.strip().splitlines() if l.startswith("{")][-1])


@pytest.fixture(scope="module")
def built(mnv):
    mnv.build_library()
    assert os.path.exists(mnv.HEADLESS_BIN)
    return mnv


@pytest.mark.parametrize("fmt,compressed", [("SH9", False), ("SH9", True), ("RGBA", False), ("SH4", True)])
def test_loader_matches_numpy(built, tmp_path, fmt, compressed):
    mnv = built
    tree = mnv.synth.make_tree(depth=5, data_format=fmt)
    path = tmp_path / "tree.npz"
    tree.save_npz(str(path), compressed=compressed)
    r = run(mnv, path, "--selftest-load")
    assert r.returncode == 0, r.stderr
    j = last_json(r.stdout)
    assert (j["N"], j["data_dim"], j["format"], j["capacity"]) == (2, tree.data_dim, fmt, tree.capacity)
    assert j["basis_dim"] == tree.basis_dim
    assert j["scale"] == [0.5, 0.5, 0.5] and j["offset"] == [0.5, 0.5, 0.5]
    assert j["child"] == mnv.bytes_checksum(tree.child.astype(np.int32))
    assert j["parent"] == mnv.bytes_checksum(tree.parent.astype(np.int32))
    assert j["data"] == mnv.bytes_checksum(tree.data.view(np.uint16))
    assert j["sample_counts_all_8"] is True  # n3tree.cpp:191-193
    assert j["pack"] == 45 and j["unpack"] == [5, 1, 0, 1]  # include/n3tree/n3tree.hpp:60-66
    assert f"Data format {fmt}, data size: {tree.capacity}" in r.stdout  # n3tree.cpp:203-204


def test_loader_invradius_scalar_and_extra_keys(built, tmp_path):
    """svox writes `invradius` (f64 scalar) in older files and several keys the viewer ignores
    (n3tree.cpp:46-52)."""
    mnv = built
    tree = mnv.synth.make_tree(depth=4)
    cap = tree.capacity
    path = tmp_path / "old.npz"
    np.savez(path, data_dim=np.int64(tree.data_dim), data_format=np.array("SH9"), invradius=np.float64(0.25),
             offset=np.array([0.5, 0.25, 0.125], np.float32), child=tree.child.reshape(cap, 2, 2, 2),
             parent_depth=np.stack([tree.parent, tree.depth], 1).astype(np.int32),
             data=tree.data.reshape(cap, 2, 2, 2, -1), n_internal=np.int64(cap), n_free=np.int64(0),
             depth_limit=np.int64(10), geom_resize_fact=np.float64(1.5), extra_data=np.zeros((0, 3), np.float32))
    r = run(mnv, path, "--selftest-load")
    assert r.returncode == 0, r.stderr
    j = last_json(r.stdout)
    assert j["scale"] == [0.25, 0.25, 0.25] and j["offset"] == [0.5, 0.25, 0.125]
    assert j["capacity"] == cap


@pytest.mark.parametrize("variant", ["stored", "deflated", "invradius_scalar", "rgba_deflated"])
def test_loader_matches_reference_loader(built, oracle, tmp_path, variant):
    """viewer::N3Tree::open (csrc/viewer/npz.cpp + n3tree.cpp) against the REFERENCE's own N3Tree::open + cnpy
    (src/n3tree/n3tree.cpp:16-205, 3rdparty/cnpy/cnpy.cpp:303-369, compiled from its sources into oracle/_ref) on
    the same file: every array the reference builds, byte for byte."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    mnv = built
    fmt = "RGBA" if variant.startswith("rgba") else "SH9"
    tree = mnv.synth.make_tree(depth=5, data_format=fmt)
    cap = tree.capacity
    path = tmp_path / f"{variant}.npz"
    if variant == "invradius_scalar":
        np.savez(path, data_dim=np.int64(tree.data_dim), data_format=np.array(fmt), invradius=np.float64(0.25),
                 offset=np.array([0.5, 0.25, 0.125], np.float32), child=tree.child.reshape(cap, 2, 2, 2),
                 parent_depth=np.stack([tree.parent, tree.depth], 1).astype(np.int32),
                 data=tree.data.reshape(cap, 2, 2, 2, -1), n_internal=np.int64(cap), n_free=np.int64(0),
                 depth_limit=np.int64(10), geom_resize_fact=np.float64(1.5), extra_data=np.zeros((0, 3), np.float32))
    else:
        tree.save_npz(str(path), compressed=variant.endswith("deflated"))
    ref = oracle.ref_load_host(str(path))
    r = run(mnv, path, "--selftest-load")
    assert r.returncode == 0, r.stderr
    j = last_json(r.stdout)
    assert (j["data_dim"], j["capacity"], j["format"], j["basis_dim"]) == \
        (ref["data_dim"], ref["capacity"], ref["format"], ref["basis_dim"])
    assert j["scale"] == [float(v) for v in ref["scale"]] and j["offset"] == [float(v) for v in ref["offset"]]
    assert j["child"] == mnv.bytes_checksum(ref["child"])
    assert j["parent"] == mnv.bytes_checksum(ref["parent"])
    assert j["data"] == mnv.bytes_checksum(ref["data"])
    assert j["sample_counts_all_8"] is True and (ref["sample_counts"] == 8).all()
    # and both equal what numpy wrote
    assert np.array_equal(ref["child"], tree.child) and np.array_equal(ref["data"], tree.data.view(np.uint16))


@pytest.mark.parametrize("n_retain", [0, 1, 3])
def test_loader_vq_compressed_tree(built, tmp_path, n_retain):
    """quant_colors / quant_map / data_retained / sigma (n3tree.cpp:109-175): basis functions
    [n_retain, n_basis) through per-basis 65536-entry codebooks, the first n_retain stored plainly,
    destination index channel * n_basis + basis (the reference's own write index drops `+ basis`)."""
    mnv = built
    tree = mnv.synth.make_tree(depth=4)
    cap, nb = tree.capacity, 9
    n_q = nb - n_retain
    rng = np.random.default_rng(11 + n_retain)
    book = rng.standard_normal((n_q, 65536, 3)).astype(np.float16)
    qmap = rng.integers(0, 65536, (n_q, cap, 2, 2, 2), dtype=np.uint16)
    retained = rng.standard_normal((n_retain, cap, 2, 2, 2, 3)).astype(np.float16)
    sigma = tree.data[..., -1].reshape(cap, 2, 2, 2)
    want = np.zeros((cap, 8, 28), np.float16)
    for b in range(n_q):
        col = book[b][qmap[b].reshape(cap, 8)]            # [cap, 8, 3]
        for ch in range(3):
            want[:, :, ch * nb + n_retain + b] = col[..., ch]
    for b in range(n_retain):
        for ch in range(3):
            want[:, :, ch * nb + b] = retained[b].reshape(cap, 8, 3)[..., ch]
    want[:, :, 27] = sigma.reshape(cap, 8)
    arrays = dict(data_dim=np.int64(28), data_format=np.array("SH9"), invradius3=tree.scale, offset=tree.offset,
                  child=tree.child.reshape(cap, 2, 2, 2), parent_depth=np.stack([tree.parent, tree.depth], 1).astype(np.int32),
                  quant_colors=book, quant_map=qmap, sigma=sigma)
    if n_retain:
        arrays["data_retained"] = retained
    path = tmp_path / "vq.npz"
    np.savez_compressed(path, **arrays)
    r = run(mnv, path, "--selftest-load")
    assert r.returncode == 0, r.stderr
    assert "Decoding quantized colors" in r.stdout
    j = last_json(r.stdout)
    assert j["capacity"] == cap and j["data_dim"] == 28
    assert j["data"] == mnv.bytes_checksum(want.view(np.uint16))
    # schema violations
    bad = dict(arrays, quant_colors=book.astype(np.float32))
    np.savez(tmp_path / "bad.npz", **bad)
    r = run(mnv, tmp_path / "bad.npz", "--selftest-load")
    assert r.returncode == 1 and "half precision" in r.stderr
    bad = dict(arrays, quant_map=qmap[:-1])
    np.savez(tmp_path / "bad2.npz", **bad)
    r = run(mnv, tmp_path / "bad2.npz", "--selftest-load")
    assert r.returncode == 1


@pytest.mark.parametrize("fmt", ["SH9", "RGBA"])
def test_tree_writer_round_trip(built, tmp_path, fmt):
    """viewer::N3Tree::save (the reference has no writer): numpy reads the file back array for array, zipfile
    accepts the archive, and the C++ loader reads its own output."""
    mnv = built
    tree = mnv.synth.make_tree(depth=5, data_format=fmt)
    src, dst = tmp_path / "a.npz", tmp_path / "b.npz"
    tree.save_npz(str(src), compressed=True)
    r = run(mnv, src, "--selftest-resave", dst)
    assert r.returncode == 0, r.stderr
    with zipfile.ZipFile(dst) as z:
        assert z.testzip() is None  # CRCs
        assert sorted(z.namelist()) == sorted(k + ".npy" for k in
                                              ("data_dim", "data_format", "invradius3", "offset", "child", "parent_depth", "data"))
    back = mnv.HostTree.load_npz(str(dst))
    assert back.data_format == fmt and back.data_dim == tree.data_dim
    assert np.array_equal(back.child, tree.child) and np.array_equal(back.parent, tree.parent)
    assert np.array_equal(back.depth, tree.depth)  # recomputed from the links
    assert np.array_equal(back.data.view(np.uint16), tree.data.view(np.uint16))
    assert np.array_equal(back.scale, tree.scale) and np.array_equal(back.offset, tree.offset)
    a, b = run(mnv, src, "--selftest-load"), run(mnv, dst, "--selftest-load")
    assert a.returncode == 0 and b.returncode == 0 and last_json(a.stdout) == last_json(b.stdout)


def test_loader_errors(built, tmp_path):
    mnv = built
    r = run(mnv, tmp_path / "nope.npz", "--selftest-load")  # n3tree.cpp:19-22: message, empty tree
    assert r.returncode == 3 and "file does not exist" in r.stdout
    tree = mnv.synth.make_tree(depth=3)
    cap = tree.capacity
    base = dict(data_dim=np.int64(28), data_format=np.array("SH9"), invradius3=tree.scale, offset=tree.offset,
                child=tree.child.reshape(cap, 2, 2, 2), parent_depth=np.stack([tree.parent, tree.depth], 1),
                data=tree.data.reshape(cap, 2, 2, 2, -1))
    bad = dict(base, data=base["data"].astype(np.float32))  # n3tree.cpp:180 "data must be half"
    np.savez(tmp_path / "f32.npz", **bad)
    r = run(mnv, tmp_path / "f32.npz", "--selftest-load")
    assert r.returncode == 1 and "half precision" in r.stderr
    bad = dict(base, parent_depth=base["parent_depth"][:-1])  # :196 sizes not aligned
    np.savez(tmp_path / "mis.npz", **bad)
    r = run(mnv, tmp_path / "mis.npz", "--selftest-load")
    assert r.returncode == 1 and "not aligned" in r.stderr
    bad = dict(base)
    del bad["child"]
    np.savez(tmp_path / "nochild.npz", **bad)
    r = run(mnv, tmp_path / "nochild.npz", "--selftest-load")
    assert r.returncode == 1 and "child" in r.stderr
    (tmp_path / "garbage.npz").write_bytes(b"not a zip file at all" * 10)
    r = run(mnv, tmp_path / "garbage.npz", "--selftest-load")
    assert r.returncode == 1
    # a truncated archive must fail cleanly, not crash
    blob = (tmp_path / "f32.npz").read_bytes()
    (tmp_path / "trunc.npz").write_bytes(blob[: len(blob) // 2])
    r = run(mnv, tmp_path / "trunc.npz", "--selftest-load")
    assert r.returncode == 1


def test_loader_plain_zip_without_zip64(built, tmp_path):
    """np.savez forces ZIP64 local headers; archives re-packed by other tools do not have them."""
    mnv = built
    tree = mnv.synth.make_tree(depth=4)
    src = tmp_path / "a.npz"
    tree.save_npz(str(src))
    dst = tmp_path / "b.npz"
    with zipfile.ZipFile(src) as zi, zipfile.ZipFile(dst, "w", zipfile.ZIP_DEFLATED) as zo:
        for name in zi.namelist():
            zo.writestr(name, zi.read(name))
    a, b = run(mnv, src, "--selftest-load"), run(mnv, dst, "--selftest-load")
    assert a.returncode == 0 and b.returncode == 0
    assert last_json(a.stdout) == last_json(b.stdout)


def wireframe_np(tree, max_depth):
    """gen_wireframe restated (n3tree.cpp:248-336): 12 edges (24 vertices of 9 floats) per leaf or
    per node cut at max_depth, children visited in (i, j, k) order."""
    out = []

    def box(bb):
        for i in range(2):
            for j in range(2):
                for (a, b, c) in ((0, i, j), (1, i, j), (i, 0, j), (i, 1, j), (i, j, 0), (i, j, 1)):
                    out.extend([bb[a * 3], bb[b * 3 + 1], bb[c * 3 + 2], 0, 0, 0, 0, 0, 1])

    def rec(node, xi, yi, zi, depth, grid):
        cnt = 0
        for i in range(xi * 2, xi * 2 + 2):
            for j in range(yi * 2, yi * 2 + 2):
                for k in range(zi * 2, zi * 2 + 2):
                    c = int(tree.child[node, cnt])
                    if c == 0 or depth >= max_depth:
                        ijk = (i, j, k)
                        lo = [np.float32((np.float32(ijk[a]) / np.float32(grid) - tree.offset[a]) / tree.scale[a])
                              for a in range(3)]
                        hi = [np.float32((np.float32(ijk[a] + 1) / np.float32(grid) - tree.offset[a]) / tree.scale[a])
                              for a in range(3)]
                        box(lo + hi)
                    else:
                        rec(node + c, i, j, k, depth + 1, grid * 2)
                    cnt += 1

    rec(0, 0, 0, 0, 0, 2)
    return np.asarray(out, np.float32)


@pytest.mark.parametrize("max_depth", [0, 2, 100])
def test_wireframe(built, tmp_path, max_depth):
    mnv = built
    tree = mnv.synth.make_tree(depth=4)
    path = tmp_path / "t.npz"
    tree.save_npz(str(path))
    r = run(mnv, path, "--selftest-wireframe", max_depth)
    assert r.returncode == 0, r.stderr
    j = last_json(r.stdout)
    want = wireframe_np(tree, max_depth)
    assert j["floats"] == want.size
    assert j["hash"] == mnv.bytes_checksum(want)


def camera_np(width, height, fx, back, up, center):
    """Camera::_update (src/camera.cpp:54-111) in float64 -> (transform 12, K 16, w2c 16)."""
    back = back / np.linalg.norm(back)
    right = np.cross(up, back)
    right /= np.linalg.norm(right)
    vup = np.cross(back, right)
    T = np.stack([right, vup, back, center])  # columns
    K = np.zeros((4, 4))
    K[0, 0] = fx / (0.5 * width)
    K[1, 1] = -fx / (0.5 * height)
    K[2, 2] = K[2, 3] = -1.0
    K[3, 2] = -2e-3
    R = T[:3].T  # world <- camera
    w2c = np.eye(4)
    w2c[:3, :3] = R.T
    w2c[:3, 3] = -R.T @ center
    return T.ravel(), K.ravel(), w2c.T.ravel()  # column-major like glm


def test_camera_defaults_and_update(built):
    mnv = built
    r = run(mnv, "--selftest-camera", "--width", 800, "--height", 600)
    assert r.returncode == 0, r.stderr
    steps = [json.loads(l) for l in r.stdout.strip().splitlines()]
    assert len(steps) == 3
    s0 = steps[0]
    assert (s0["width"], s0["height"], s0["fx"], s0["fy"], s0["cx"], s0["cy"]) == (800, 600, 1111, 1111, 400, 300)
    T, K, w2c = camera_np(800, 600, 1111.0, np.array([-0.7071068, 0, 0.7071068]), np.array([0, 0, 1.0]),
                          np.array([-3.55, 0, 3.55]))  # defaults of src/camera.cpp:41-44
    assert np.allclose(s0["transform"], T, atol=1e-6)
    assert np.allclose(s0["K"], K, atol=1e-6)
    assert np.allclose(s0["w2c"], w2c, atol=1e-5)
    # after an orbit drag the pose is still orthonormal and the eye did not move; a pan + move does
    for s in steps[1:]:
        M = np.asarray(s["transform"]).reshape(4, 3)
        assert np.allclose(M[:3] @ M[:3].T, np.eye(3), atol=1e-5)
        W = np.asarray(s["w2c"]).reshape(4, 4).T
        C = np.eye(4)
        C[:3, :3] = M[:3].T
        C[:3, 3] = M[3]
        assert np.allclose(W @ C, np.eye(4), atol=1e-5)
    assert np.allclose(steps[1]["transform"][9:], [-3.55, 0, 3.55], atol=1e-6)
    assert not np.allclose(steps[1]["transform"][:9], steps[0]["transform"][:9], atol=1e-3)
    assert not np.allclose(steps[2]["transform"][9:], [-3.55, 0, 3.55], atol=1e-3)


def test_camera_matches_reference_golden(built):
    """tests/golden/camera_trace.json: the same script run on the reference's own viewer::Camera
    (oracle/ref_driver.cpp ref_camera_trace, generated on the GPU box by tests/test_viewer_gpu.py)."""
    mnv = built
    p = os.path.join(os.path.dirname(__file__), "golden", "camera_trace.json")
    if not os.path.exists(p):
        pytest.skip("golden camera trace not generated yet")
    g = json.load(open(p))
    r = run(mnv, "--selftest-camera", "--width", g["width"], "--height", g["height"])
    steps = [json.loads(l) for l in r.stdout.strip().splitlines()]
    for s, ref in zip(steps, g["steps"]):
        got = np.asarray(s["transform"] + s["K"] + s["w2c"] + [s["fx"], s["fy"], s["cx"], s["cy"]], np.float32)
        assert np.allclose(got, np.asarray(ref, np.float32), rtol=2e-6, atol=2e-6), np.abs(got - ref).max()


def test_gl_presentation_shim_type_checks(built):
    """viewer::GlPresenter (csrc/viewer/gl_interop.cpp; reference cuda_renderer.cpp:43-66,70-95,156-162,383-458) is
    built only with -DMNV_WITH_GL; without GL development packages it is compiled against declarations-only headers."""
    csrc = os.path.join(os.path.dirname(built.LIB_PATH), "csrc")
    r = subprocess.run(["make", "-C", csrc, "gl-check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
