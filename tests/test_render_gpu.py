"""GPU parity tests (pytest -m gpu, run on the B200 box) — all through the C-ABI.

  1. native CUDA path vs the committed golden vectors (outputs of the
     reference's own kernel): bit-exact pixels, visit hashes/counts/logs and
     split / re-sample candidates;
  2. native vs the CPU oracle on seeded inputs (pixels <= 1/255, >= 50 dB,
     visit sequences equal up to the expf-ulp prefix rule);
  3. full BASELINE size (1920x1080, depth-10 tree): size-independent properties
     — tile-partition union == whole frame, logged == unlogged, idempotence,
     host-buffer call == device call — and, when oracle/_ref is present, bit
     equality with the reference kernel itself.
"""
import os

import numpy as np
import pytest

from helpers import golden_names, load_golden, psnr, sequences_prefix_equal

pytestmark = pytest.mark.gpu
LOG_CAP = 96


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "these tests need a CUDA device"
    return torch


@pytest.mark.parametrize("name", golden_names())
def test_native_matches_reference_golden_bit_exact(name, mnv, torch_cuda):
    torch = torch_cuda
    tree, cam, okw, ref = load_golden(name)
    opt = mnv.default_options(**okw)
    dt = mnv.DeviceTree(tree)
    P = cam["width"] * cam["height"]
    m = dt.render_logged(cam, opt, log_cap=LOG_CAP)
    ts = torch.empty((P, 3), device="cuda")
    tp = torch.empty((P, 3), device="cuda")
    img = dt.render(cam, opt, to_split=ts, to_sample=tp).cpu().numpy()
    assert np.array_equal(img, ref["rgba"]), "pixels differ from the reference kernel"
    assert np.array_equal(m["rgba"], ref["rgba"])
    assert np.array_equal(m["hash"], ref["hash"]), "leaf-visit sequence hash differs"
    assert np.array_equal(m["count"], ref["count"])
    assert np.array_equal(m["log"], ref["log"])
    assert np.array_equal(ts.cpu().numpy(), ref["to_split"])
    assert np.array_equal(tp.cpu().numpy(), ref["to_sample"])
    dt.close()


@pytest.mark.parametrize("fmt,depth", [("SH9", 6), ("RGBA", 5), ("SH4", 5), ("SH16", 4), ("SH25", 4), ("SH1", 5)])
def test_native_matches_oracle_seeded(fmt, depth, mnv, oracle, torch_cuda):
    tree = mnv.synth.make_tree(depth=depth, data_format=fmt, seed=7)
    dt = mnv.DeviceTree(tree)
    for pose in (1, 6, 11):
        cam = mnv.synth.default_camera(160, 90, pose=pose)
        kw = dict(background_brightness=0.25, basis_minmax=[0, max(tree.basis_dim - 1, 0)])
        o = oracle.render_voxels(tree, cam, oracle.default_options(**kw), trackers=True, log_cap=LOG_CAP)
        m = dt.render_logged(cam, mnv.default_options(**kw), log_cap=LOG_CAP)
        d = np.abs(o["rgba"].astype(int) - m["rgba"].astype(int))
        assert d.max() <= 1 and psnr(o["rgba"], m["rgba"]) >= 50.0  # 1/255 max-abs, 50 dB
        same = (o["hash"] == m["hash"]) & (o["count"] == m["count"])
        bad = np.nonzero(~same)[0]
        assert bad.size <= max(2, same.size // 500)
        for ray in bad:
            assert sequences_prefix_equal(o["log"], o["count"], m["log"], m["count"], ray)
    dt.close()


def test_point_query_matches_integer_descent(mnv, torch_cuda):
    tree = mnv.synth.make_tree(depth=8)
    dt = mnv.DeviceTree(tree)
    pts = np.random.default_rng(2).random((1_000_000, 3)).astype(np.float32)
    pts[:4] = [[0, 0, 0], [1, 1, 1], [-1, 2, 0.5], [1 - 1e-7, 0.5, 0.25]]
    q = dt.query_points(pts).cpu().numpy()
    assert np.array_equal(q, mnv.synth.brute_force_query(tree, pts))
    dt.close()


def test_device_tree_roundtrip_and_ragged_sizes(mnv, torch_cuda):
    tree = mnv.synth.make_tree(depth=5)
    sc = np.random.default_rng(0).integers(0, 300, (tree.capacity, 8)).astype(np.int16)
    dt = mnv.DeviceTree(tree, max_capacity=tree.capacity + 1000, sample_counts=sc)
    data, child, parent, counts = dt.download()
    assert np.array_equal(child, tree.child) and np.array_equal(parent, tree.parent)
    assert np.array_equal(data.view(np.uint16), tree.data.view(np.uint16))
    assert np.array_equal(counts, sc)
    assert dt.capacity == tree.capacity and dt.max_capacity == tree.capacity + 1000
    opt = mnv.default_options(background_brightness=0.0)
    # ragged frame sizes (not multiples of the 16x8 block tile), 1x1 frame
    for (w, h) in ((1, 1), (17, 9), (33, 5), (250, 131)):
        cam = mnv.synth.default_camera(w, h, pose=3)
        img = dt.render(cam, opt).cpu().numpy()
        assert img.shape == (h, w, 4) and (img[..., 3] == 255).all()
    dt.close()


@pytest.fixture(scope="module")
def big(mnv, torch_cuda):
    """Config 2 of BASELINE.json: depth-10 SH9 octree, 1920x1080."""
    tree = mnv.synth.make_tree(depth=10)
    dt = mnv.DeviceTree(tree)
    cam = mnv.synth.default_camera(1920, 1080)
    opt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8])
    yield tree, dt, cam, opt
    dt.close()


def test_full_size_properties(big, mnv, torch_cuda):
    torch = torch_cuda
    tree, dt, cam, opt = big
    full = dt.render(cam, opt).cpu().numpy()
    # idempotence / determinism
    assert np.array_equal(full, dt.render(cam, opt).cpu().numpy())
    # logged kernel == production kernel, and the per-ray counters are consistent
    m = dt.render_logged(cam, opt)
    assert np.array_equal(m["rgba"], full)
    assert (m["shaded"] <= m["count"]).all() and m["count"].max() > 0
    # tile partition (multi-GPU split) : union of N disjoint tile sets == whole frame
    for n in (2, 8):
        out = torch.zeros((1080, 1920, 4), dtype=torch.uint8, device="cuda")
        for r in range(n):
            dt.render_tiles(cam, opt, out, 64, 64, n, r)
        assert np.array_equal(out.cpu().numpy(), full)
    # the host-buffer frame call == the device call, and its statistics add up
    host, st = dt.render_frame_host(cam, opt, stats=True)
    assert np.array_equal(host, full)
    assert st["rays"] == 1920 * 1080 and st["visits"] == int(m["count"].sum())
    assert st["shaded_visits"] == int(m["shaded"].sum())
    # trackers do not change the image
    ts = torch.empty((1920 * 1080, 3), device="cuda")
    tp = torch.empty((1920 * 1080, 3), device="cuda")
    assert np.array_equal(dt.render(cam, opt, to_split=ts, to_sample=tp).cpu().numpy(), full)
    tsn = ts.cpu().numpy()
    assert (tsn[:, 0] >= 1).all() and (tsn[:, 1] < tree.capacity).all()


def test_full_size_matches_reference_kernel(big, mnv, oracle, torch_cuda, tmp_path):
    """Bit equality with the reference's own kernel at BASELINE size (needs oracle/_ref)."""
    if not (oracle.ref_available() and oracle.ref_available(instr=True)):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    torch = torch_cuda
    tree, dt, cam, opt = big
    npz = str(tmp_path / "tree.npz")
    tree.save_npz(npz)
    oopt = oracle.default_options(background_brightness=0.0, basis_minmax=[0, 8])
    ref = oracle.RefRenderer(npz)
    r = ref.render(cam, oopt)
    ts = torch.empty((1920 * 1080, 3), device="cuda")
    tp = torch.empty((1920 * 1080, 3), device="cuda")
    img = dt.render(cam, opt, to_split=ts, to_sample=tp).cpu().numpy()
    assert np.array_equal(img, r["rgba"])
    assert np.array_equal(ts.cpu().numpy(), r["to_split"])
    assert np.array_equal(tp.cpu().numpy(), r["to_sample"])
    ref.close()
    refi = oracle.RefRenderer(npz, instr=True)
    ri = refi.render_logged(cam, oopt)
    m = dt.render_logged(cam, opt)
    assert np.array_equal(ri["hash"], m["hash"]) and np.array_equal(ri["count"], m["count"])
    refi.close()


@pytest.mark.parametrize("w,h", [(1, 1), (7, 5), (33, 17), (127, 65), (250, 141)])
def test_ragged_frame_sizes_match_the_reference_kernel(mnv, oracle, tmp_path, w, h):
    """Frames that are not multiples of the 16x8 CTA tile / 8x4 warp tile, down to a single pixel, with a principal
    point off the half-pixel grid: pixels and candidates equal to the reference kernel's (oracle/_ref), or to
    the CPU port within one LSB when the reference build is absent."""
    import torch

    tree = mnv.synth.make_tree(depth=6)
    cam = mnv.synth.default_camera(w, h, pose=6)
    cam["cx"], cam["cy"] = w * 0.47 + 0.3, h * 0.52 - 0.2
    okw = dict(background_brightness=0.25, basis_minmax=[0, 8])
    dt = mnv.DeviceTree(tree)
    P = w * h
    ts, tp = torch.empty((P, 3), device="cuda"), torch.empty((P, 3), device="cuda")
    got = dt.render(cam, mnv.default_options(**okw), to_split=ts, to_sample=tp).cpu().numpy()
    assert got.shape == (h, w, 4) and (got[..., 3] == 255).all()
    if oracle.ref_available():
        npz = str(tmp_path / "t.npz")
        tree.save_npz(npz)
        ref = oracle.RefRenderer(npz)
        r = ref.render(cam, oracle.default_options(**okw))
        assert np.array_equal(got, r["rgba"])
        assert np.array_equal(ts.cpu().numpy(), r["to_split"]) and np.array_equal(tp.cpu().numpy(), r["to_sample"])
        ref.close()
    else:
        o = oracle.render_voxels(tree, cam, oracle.default_options(**okw))
        assert np.abs(got.astype(int) - o["rgba"].astype(int)).max() <= 1
    dt.close()


def test_tile_launch_order_does_not_change_the_frame(mnv):
    """mnv_tree_set_tile_order: any permutation of the CTA tiles gives the same pixels and candidates."""
    import torch

    tree = mnv.synth.make_tree(depth=6)
    dt = mnv.DeviceTree(tree)
    w, h = 330, 150
    cam = mnv.synth.default_camera(w, h, pose=7)
    opt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8])
    P = w * h
    def frame():
        ts, tp = torch.empty((P, 3), device="cuda"), torch.empty((P, 3), device="cuda")
        img = dt.render(cam, opt, to_split=ts, to_sample=tp)
        return img.cpu().numpy(), ts.cpu().numpy(), tp.cpu().numpy()
    want = frame()
    n_tiles = ((w + 15) // 16) * ((h + 7) // 8)
    for seed in (0, 1):
        order = torch.from_numpy(np.random.default_rng(seed).permutation(n_tiles).astype(np.int32)).cuda()
        dt.set_tile_order(order)
        got = frame()
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
    dt.set_tile_order(torch.arange(7, dtype=torch.int32, device="cuda"))  # wrong length: ignored
    assert np.array_equal(frame()[0], want[0])
    dt.set_tile_order(None)
    dt.close()
