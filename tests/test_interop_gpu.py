"""GPU parity of the two modes of render_voxels the octree goldens do not exercise, against the reference's own
kernel (oracle/_ref): the viewer's presentation path — offscreen = false, compositing over the colour surface GL
already drew and clipping against the R32F mesh-depth surface (renderer_kernel.cu:259-264,277-280,225-229;
cuda_renderer.cpp:141-142) — and visit tracking (`track_visit`, the atomicCAS marks of rt_core.cuh:133-135 that
Impl::prune_tree consumes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scene(mnv, tmp_path, depth=6):
    tree = mnv.synth.make_tree(depth=depth)
    npz = str(tmp_path / "t.npz")
    tree.save_npz(npz)
    return tree, npz


@pytest.mark.parametrize("pose", [0, 5])
def test_interop_surfaces_match_the_reference_kernel(mnv, oracle, tmp_path, pose):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    tree, npz = _scene(mnv, tmp_path)
    w, h = 320, 180
    cam = mnv.synth.default_camera(w, h, pose=pose)
    rng = np.random.default_rng(pose)
    prior = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)   # what the mesh pass left in the colour buffer
    prior[..., 3] = 255
    # mesh depth: a third of the pixels have geometry in front of / inside the volume, the rest are "infinitely" far
    depth = np.full((h, w), 1e9, np.float32)
    near = rng.random((h, w)) < 0.33
    depth[near] = rng.uniform(0.5, 4.0, near.sum()).astype(np.float32)
    okw = dict(background_brightness=1.0, basis_minmax=[0, 8])
    ref = oracle.RefRenderer(npz)
    want, _ = ref.render_interop(cam, oracle.default_options(**okw), prior, depth)
    dt = mnv.DeviceTree(tree)
    got = dt.render_interop(cam, mnv.default_options(**okw), prior, depth)
    assert np.array_equal(got, want), np.abs(got.astype(int) - want.astype(int)).max()
    # the depth surface really clips: with no geometry the frame differs where `near` cut rays short
    free = dt.render_interop(cam, mnv.default_options(**okw), prior, np.full((h, w), 1e9, np.float32))
    assert (free != got).any() and np.array_equal(free[~near], got[~near])
    ref.close()
    dt.close()


def test_visit_tracking_matches_the_reference_kernel(mnv, oracle, tmp_path):
    import torch

    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    tree, npz = _scene(mnv, tmp_path, depth=7)
    w, h = 256, 144
    cam = mnv.synth.default_camera(w, h, pose=3)
    okw = dict(background_brightness=0.0, basis_minmax=[0, 8])
    cap = tree.capacity
    prior = np.zeros((h, w, 4), np.uint8)
    depth = np.full((h, w), 1e9, np.float32)
    ref = oracle.RefRenderer(npz, max_capacity=cap)
    want_img, want_vis = ref.render_interop(cam, oracle.default_options(**okw), prior, depth, track_visit=True,
                                            max_capacity=cap)
    dt = mnv.DeviceTree(tree, max_capacity=cap)
    visited = torch.zeros(cap, dtype=torch.int32, device="cuda")
    visited[0] = 1
    P = w * h
    ts, tp = torch.empty((P, 3), device="cuda"), torch.empty((P, 3), device="cuda")
    got_img = dt.render_interop(cam, mnv.default_options(**okw), prior, depth, track_visit=True, visited=visited,
                                to_split=ts, to_sample=tp)
    got_vis = visited.cpu().numpy()
    assert np.array_equal(got_img, want_img)
    assert np.array_equal(got_vis != 0, want_vis != 0)
    assert 1 < (got_vis != 0).sum() < cap  # some, not all, nodes are seen from this pose
    # what prune_tree would drop is the same set
    num_unvisited = int((got_vis[:cap] == 0).sum())
    assert num_unvisited == int((want_vis[:cap] == 0).sum())
    ref.close()
    dt.close()
