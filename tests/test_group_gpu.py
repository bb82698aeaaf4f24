"""GPU tests of the replica group (csrc/mnv_group.cu): the image-tile multi-GPU mode driven from one process, with
NVLink peer copies into the display GPU, and dynamic refinement across the group.  On a one-GPU box the replicas
share the device (the partition, the gather, the record exchanges and the orderings are the same code); with 2+
GPUs the group spans them."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GRID, MINP, RNG = [2, 4], [-1.0, -1.0, -1.0], [2.0, 2.0, 2.0]


def _devices(n):
    import torch

    have = torch.cuda.device_count()
    return [i % have for i in range(n)]


@pytest.mark.parametrize("n,size,band", [(2, (320, 180), 8), (3, (333, 187), 16), (8, (640, 360), 8)])
def test_group_frame_equals_one_gpu_frame(n, size, band, mnv):
    import torch

    W, H = size
    tree = mnv.synth.make_tree(depth=7)
    opt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8])
    dt = mnv.DeviceTree(tree)
    grp = mnv.ReplicaGroup(tree, _devices(n))
    for pose in (0, 5):
        cam = mnv.synth.default_camera(W, H, pose=pose)
        want = dt.render(cam, opt).cpu().numpy()
        got = grp.render_frame_host(cam, opt, band_rows=band)
        assert np.array_equal(got, want)
        # asynchronous form into a device buffer on devices[0]
        out = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda:0")
        grp.render_frame(cam, opt, out, band_rows=band)
        grp.synchronize()
        assert np.array_equal(out.cpu().numpy(), want)
    grp.close()
    dt.close()


def test_group_refinement_matches_single_replica(mnv):
    """Refinement over 1 and over 3 replicas: the same leaves are split, the same payloads committed — every replica's
    tree and the frames after refinement are identical, bit for bit."""
    tree = mnv.synth.make_tree(depth=6)
    subs = [mnv.synth.make_mlp_weights(seed=3 + i) for i in range(8)]
    opt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8], use_splitting=True, appearance_embedding=0,
                              split_batch_size=512)
    cams = [mnv.synth.default_camera(384, 216, pose=p) for p in (1, 2, 9)]
    results = []
    for n in (1, 3):
        dev = _devices(n)
        grp = mnv.ReplicaGroup(tree, dev, max_capacity=tree.capacity + 4 * 512)
        models = [mnv.MlpModel(subs, grid_dim=GRID, min_position=MINP, max_position=[1.0] * 3, device=d) for d in dev]
        added, frames = 0, []
        for cam in cams:
            img, k = grp.refine_frame(models, cam, opt, GRID, MINP, RNG, seed=77)
            added += k
            frames.append(img)
        final = grp.render_frame_host(cams[0], mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8]))
        trees = [grp.replica(i).download() for i in range(n)]
        caps = [grp.replica(i).capacity for i in range(n)]
        results.append((added, frames, final, trees, caps))
        for m in models:
            m.close()
        grp.close()
    (a1, f1, fin1, t1, c1), (a3, f3, fin3, t3, c3) = results
    assert a1 == a3 > 0 and c1[0] == tree.capacity + a1 and all(c == c1[0] for c in c3)
    for x, y in zip(f1, f3):
        assert np.array_equal(x, y)
    assert np.array_equal(fin1, fin3)
    for rep in t3:
        for x, y in zip(rep, t1[0]):
            assert np.array_equal(x, y)
