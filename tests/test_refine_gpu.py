"""GPU parity tests of dynamic refinement through the C-ABI: native
kernels vs the reference's own kernels (oracle/_ref) and vs the numpy restatement."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GRID, MINP, RNG = [2, 4], [-1.0, -1.0, -1.0], [2.0, 2.0, 2.0]


def _pick_leaves(tree, n, seed):
    leaves = np.argwhere(tree.child == 0)
    rng = np.random.default_rng(seed)
    return leaves[rng.choice(len(leaves), n, replace=False)].astype(np.int32)


def _assert_renders_like_a_fresh_build(mnv, dt, tree):
    """The march reads the device planes refinement edits in place (cell words, payload records, sample counts): after
    any refinement step they must describe the same tree as a fresh upload of the downloaded arrays — render both."""
    import torch

    data, child, parent, counts = dt.download()
    t2 = mnv.HostTree(N=2, data_dim=tree.data_dim, data_format=tree.data_format, child=child, parent=parent,
                      depth=np.zeros(child.shape[0], np.int32), data=data, scale=tree.scale, offset=tree.offset)
    dt2 = mnv.DeviceTree(t2, sample_counts=counts)
    opt = mnv.default_options(background_brightness=0.0, max_sample_count=12)
    for pose in (0, 3, 9):
        cam = mnv.synth.default_camera(160, 90, pose=pose)
        P = 160 * 90
        outs = []
        for d in (dt, dt2):
            ts, tp = torch.empty((P, 3), device="cuda"), torch.empty((P, 3), device="cuda")
            img = d.render(cam, opt, to_split=ts, to_sample=tp)
            outs.append((img.cpu().numpy(), ts.cpu().numpy(), tp.cpu().numpy()))
        for a, b in zip(*outs):
            assert np.array_equal(a, b)
    dt2.close()


@pytest.mark.parametrize("okw", [dict(), dict(need_viewdir=True, appearance_embedding=1), dict(appearance_embedding=0)])
def test_add_children_commit_and_generate(okw, mnv, oracle, tmp_path):
    import torch
    from oracle import refine_np as R

    tree = mnv.synth.make_tree(depth=5)
    cap = tree.capacity
    n, c = 37, 8
    mopt, oopt = mnv.default_options(**okw), oracle.default_options(**okw)
    rd = 3 + 3 * bool(okw.get("need_viewdir")) + ("appearance_embedding" in okw)
    parents = _pick_leaves(tree, n, 1)
    rand = np.random.default_rng(2).random((n * 8, c, rd)).astype(np.float32)

    dt = mnv.DeviceTree(tree, max_capacity=cap + 100)
    samples = torch.from_numpy(rand).cuda()
    cluster = torch.zeros((n * 8, c), dtype=torch.int16, device="cuda")
    dt.add_children(mopt, torch.from_numpy(parents).cuda(), samples, cluster, GRID, MINP, RNG)
    # MLP stand-in: deterministic "results" per sample row
    D = tree.data_dim
    results = torch.from_numpy(np.random.default_rng(3).standard_normal((n * 8, c, D + 1)).astype(np.float32)).cuda()
    dt.commit_children(mopt, n, results)
    torch.cuda.synchronize()
    assert dt.capacity == cap + n
    data, child, parent, counts = dt.download()

    # links == numpy restatement; old rows untouched
    child_np, parent_np = R.add_children(tree.child, tree.parent, cap, parents)
    assert np.array_equal(child, child_np) and np.array_equal(parent, parent_np)
    assert np.array_equal(data[:cap].view(np.uint16), tree.data.view(np.uint16))
    # payload of the new leaves = fp16(mean over the c results), counts = c
    want = results.cpu().numpy()[:, :, :D].mean(1).astype(np.float16).reshape(n, 8, D)
    assert np.array_equal(data[cap:].view(np.uint16), want.view(np.uint16))
    assert (counts[cap:] == c).all() and (counts[:cap] == 8).all()
    # sample geometry / cluster ids
    packed = (np.repeat(np.arange(cap, cap + n), 8) * 8 + np.tile(np.arange(8), n))
    s_np, cl_np = R.generate_samples(parent_np, tree.scale, tree.offset, packed, rand,
                                     bool(okw.get("need_viewdir")), okw.get("appearance_embedding", -1), GRID, MINP, RNG)
    s_nat, cl_nat = samples.cpu().numpy(), cluster.cpu().numpy()
    assert np.allclose(s_nat, s_np, rtol=0, atol=1e-6) and (cl_nat != cl_np).mean() < 1e-3
    # samples lie inside their voxel's bounding box in world space
    assert s_nat[..., :3].min() >= -1.0 - 1e-6 and s_nat[..., :3].max() <= 1.0 + 1e-6

    # the refined tree renders, and new leaves are reachable by the point query
    img = dt.render(mnv.synth.default_camera(64, 36), mnv.default_options(background_brightness=0.0)).cpu().numpy()
    assert (img[..., 3] == 255).all()
    pts = s_nat[:, 0, :3] * tree.scale + tree.offset
    q = dt.query_points(pts.astype(np.float32)).cpu().numpy()
    assert np.array_equal(q[:, 0] * 8 + q[:, 1], packed)
    _assert_renders_like_a_fresh_build(mnv, dt, tree)

    # generate_samples for existing leaves (A11) + running-mean update
    nodes = _pick_leaves(tree, 50, 4)
    rand2 = np.random.default_rng(5).random((50, c, rd)).astype(np.float32)
    s2 = torch.from_numpy(rand2).cuda()
    cl2 = torch.zeros((50, c), dtype=torch.int16, device="cuda")
    # leaves that were split above are internal now: restrict to still-leaf ones
    is_leaf = child[nodes[:, 0], nodes[:, 1]] == 0
    nodes = nodes[is_leaf]
    s2, cl2, rand2 = s2[: len(nodes)].contiguous(), cl2[: len(nodes)].contiguous(), rand2[: len(nodes)]
    s2.copy_(torch.from_numpy(rand2))
    dt.generate_samples(mopt, torch.from_numpy(nodes).cuda(), s2, cl2, GRID, MINP, RNG)
    res2 = torch.from_numpy(np.random.default_rng(6).standard_normal((len(nodes), c, D + 1)).astype(np.float32)).cuda()
    dt.update_samples(mopt, torch.from_numpy(nodes).cuda(), res2)
    torch.cuda.synchronize()
    data2, _, _, counts2 = dt.download()
    old = data[nodes[:, 0], nodes[:, 1]].astype(np.float32)
    new_sum = res2.cpu().numpy()[:, :, :D].sum(1)
    want2 = old + (new_sum - c * old) / 16.0
    got2 = data2[nodes[:, 0], nodes[:, 1]].astype(np.float32)
    assert np.allclose(got2, want2, rtol=2e-3, atol=2e-3)
    assert (counts2[nodes[:, 0], nodes[:, 1]] == 16).all()
    _assert_renders_like_a_fresh_build(mnv, dt, tree)

    if oracle.ref_available():
        npz = str(tmp_path / "t.npz")
        tree.save_npz(npz)
        ref = oracle.RefRenderer(npz, max_capacity=cap + 100)
        rs, rcl = ref.add_children(oopt, parents, rand, GRID, MINP, RNG)
        assert np.array_equal(s_nat, rs), np.abs(s_nat - rs).max()   # bit-exact vs the reference kernel
        assert np.array_equal(cl_nat, rcl)
        _, rchild, rparent, _, _ = ref.download()
        assert np.array_equal(rchild, child) and np.array_equal(rparent, parent)
        rs2, rcl2 = ref.generate_samples(oopt, nodes, rand2, GRID, MINP, RNG)
        if okw.get("need_viewdir") or "appearance_embedding" in okw:
            # reference bug only concerns get_more_samples' 3-column buffer; the
            # kernel itself is driven here with a consistent rand_dim
            pass
        assert np.array_equal(s2.cpu().numpy(), rs2) and np.array_equal(cl2.cpu().numpy(), rcl2)
        ref.close()
    dt.close()


def test_prune_matches_reference_and_numpy(mnv, oracle, tmp_path):
    import torch
    from oracle import refine_np as R

    tree = mnv.synth.make_tree(depth=5)
    cap = tree.capacity
    # delete a set of subtrees: every node whose ancestor chain contains a marked node
    rng = np.random.default_rng(7)
    marked = np.zeros(cap, bool)
    marked[rng.choice(np.arange(1, cap), 25, replace=False)] = True
    to_delete = marked.copy()
    for node in range(1, cap):  # BFS order: parents precede children
        to_delete[node] |= to_delete[tree.parent[node] // 8]
    assert 0 < to_delete.sum() < cap and not to_delete[0]

    dt = mnv.DeviceTree(tree)
    cam = mnv.synth.default_camera(64, 36, pose=2)
    opt = mnv.default_options(background_brightness=0.0)
    num = dt.prune(torch.from_numpy(to_delete).cuda())
    torch.cuda.synchronize()
    assert num == int(to_delete.sum()) and dt.capacity == cap - num
    data, child, parent, counts = dt.download()
    c_np, p_np, d_np, keep = R.prune(tree.child, tree.parent, tree.data, to_delete)
    assert np.array_equal(child, c_np) and np.array_equal(parent[1:], p_np[1:])
    assert np.array_equal(data.view(np.uint16), d_np.view(np.uint16))
    # the pruned tree is a valid tree: renders, and equals rendering the numpy-pruned tree
    img = dt.render(cam, opt).cpu().numpy()
    t2 = mnv.HostTree(N=2, data_dim=tree.data_dim, data_format=tree.data_format, child=c_np, parent=p_np,
                      depth=tree.depth[keep], data=d_np, scale=tree.scale, offset=tree.offset)
    dt2 = mnv.DeviceTree(t2)
    assert np.array_equal(img, dt2.render(cam, opt).cpu().numpy())
    dt2.close()
    if oracle.ref_available():
        npz = str(tmp_path / "t.npz")
        tree.save_npz(npz)
        ref = oracle.RefRenderer(npz)
        assert ref.prune(to_delete) == num
        rdata, rchild, rparent, _, _ = ref.download()
        assert np.array_equal(rchild, child) and np.array_equal(rparent[1:], parent[1:])
        assert np.array_equal(rdata.view(np.uint16), data.view(np.uint16))
        ref.close()
    dt.close()
