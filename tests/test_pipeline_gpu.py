"""GPU tests of the device-side refinement glue: candidate selection
(the unique_dim / sort chain of Impl::expand_voxels and get_more_samples) and the per-sub-module
MLP dispatch of Impl::query_submodules, each against a plain numpy / torch restatement of the
reference's LibTorch ops."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def ref_select_split(tracker, max_n):
    """cuda_renderer.cpp:206-226 in numpy (np.unique sorts rows lexicographically like unique_dim)."""
    cand = tracker[tracker[:, 1] >= 0]
    if len(cand) == 0:
        return np.zeros((0, 2), np.int32), 0
    rows, counts = np.unique(cand, axis=0, return_counts=True)
    t = np.concatenate([-counts[:, None].astype(np.float32), rows], 1)
    t = t[t[:, 0] < -1]
    t = np.unique(t, axis=0)
    return t[:max_n, 2:].astype(np.int32), len(t)


def ref_select_sample(tracker, max_n):
    cand = tracker[tracker[:, 1] >= 0]
    rows = np.unique(cand, axis=0)
    return rows[:max_n, 1:].astype(np.int32), len(rows)


def test_candidate_selection_matches_torch_semantics(mnv):
    import torch

    tree = mnv.synth.make_tree(depth=7)
    dt = mnv.DeviceTree(tree)
    cam = mnv.synth.default_camera(480, 270, pose=3)
    opt = mnv.default_options(background_brightness=0.0, max_depth=7, max_sample_count=9)
    P = 480 * 270
    ts = torch.empty((P, 3), device="cuda")
    tp = torch.empty((P, 3), device="cuda")
    dt.render(cam, opt, to_split=ts, to_sample=tp)
    torch.cuda.synchronize()
    for max_n in (64, 4096):
        nodes, nc = mnv.select_candidates(ts, max_n, "split")
        want, wc = ref_select_split(ts.cpu().numpy(), max_n)
        assert nc == wc and wc > 0
        assert np.array_equal(nodes.cpu().numpy(), want)
        nodes, nc = mnv.select_candidates(tp, max_n, "sample")
        want, wc = ref_select_sample(tp.cpu().numpy(), max_n)
        assert nc == wc and wc > 0
        assert np.array_equal(nodes.cpu().numpy(), want)
    # no candidates at all (every ray misses / nothing splittable)
    none = torch.full((1000, 3), -1.0, device="cuda")
    nodes, nc = mnv.select_candidates(none, 16, "split")
    assert nodes.shape[0] == 0 and nc == 0
    dt.close()


def _random_tracker(rng, P, n_leaves, id_base=0, prio_max=9, none_frac=0.3):
    """Rows (priority, chunk, child) as the march kernel writes them: priority is a function of the leaf."""
    ids = id_base + rng.integers(0, n_leaves, P) * 3  # sparse leaf ids
    ids = np.where(rng.random(P) < 0.5, ids, id_base + (rng.zipf(1.6, P) % n_leaves) * 3)  # heavy hitters
    prio = (ids * 2654435761 % (prio_max + 1)).astype(np.float32)
    rows = np.stack([prio, (ids >> 3).astype(np.float32), (ids & 7).astype(np.float32)], 1).astype(np.float32)
    rows[rng.random(P) < none_frac] = -1.0
    return rows


@pytest.mark.parametrize("P,n_leaves,max_n", [(200_000, 50_000, 4096), (70_001, 300, 4192), (5000, 4000, 16384),
                                               (1_000_000, 400_000, 1), (33, 5, 8)])
def test_vote_selection_random_rows(P, n_leaves, max_n, mnv):
    """Hash-count + radix-select selection (csrc/mnv_vote.cu) against the unique_dim chain in numpy."""
    import torch

    rng = np.random.default_rng(P + max_n)
    rows = _random_tracker(rng, P, n_leaves)
    t = torch.from_numpy(rows).cuda()
    for _ in range(2):  # the table must be left clean by the first call
        nodes, nc = mnv.select_candidates(t, max_n, "split")
        want, wc = ref_select_split(rows, max_n)
        assert nc == wc
        assert np.array_equal(nodes.cpu().numpy(), want)
        nodes, nc = mnv.select_candidates(t, max_n, "sample")
        want, wc = ref_select_sample(rows, max_n)
        assert nc == wc
        assert np.array_equal(nodes.cpu().numpy(), want)


def test_vote_records_merge_equals_global_selection(mnv):
    """Multi-GPU refinement: per-rank vote records, gathered and merged, select exactly what the raw rows do."""
    import torch

    rng = np.random.default_rng(77)
    rows = _random_tracker(rng, 300_000, 40_000)
    parts = np.array_split(rows, [90_000, 90_001, 210_000])  # four uneven "ranks"
    recs = [mnv.vote_reduce(torch.from_numpy(np.ascontiguousarray(p)).cuda()) for p in parts]
    # every rank's records: unique leaves with their multiplicities
    r0 = recs[0].cpu().numpy().astype(np.int64)
    v0 = parts[0][parts[0][:, 1] >= 0]
    u, c = np.unique((v0[:, 1].astype(np.int64) * 8 + v0[:, 2].astype(np.int64)), return_counts=True)
    order = np.argsort(r0[:, 0])
    assert np.array_equal(r0[order, 0], u) and np.array_equal(r0[order, 2], c)
    pad = torch.zeros((1000, 3), dtype=torch.int32, device="cuda")  # all-gather padding: votes == 0
    gathered = torch.cat([recs[0], pad, recs[1], recs[2], pad, recs[3]])
    for kind, ref in (("split", ref_select_split), ("sample", ref_select_sample)):
        nodes, nc = mnv.select_from_votes(gathered, 4096, kind)
        want, wc = ref(rows, 4096)
        assert nc == wc and np.array_equal(nodes.cpu().numpy(), want)
    # records of the other ranks + own raw rows
    nodes, nc = mnv.select_from_votes(torch.cat(recs[1:]), 4096, "split",
                                      tracker=torch.from_numpy(np.ascontiguousarray(parts[0])).cuda())
    want, wc = ref_select_split(rows, 4096)
    assert nc == wc and np.array_equal(nodes.cpu().numpy(), want)
    with pytest.raises(mnv.MnvError) as ei:
        mnv.vote_reduce(torch.from_numpy(np.ascontiguousarray(parts[2])).cuda(), cap_records=100)
    assert ei.value.code == 7  # MNV_ERR_FULL


def test_vote_selection_node_ids_above_2_pow_24(mnv):
    """float trackers are exact only below 2^24 nodes; larger ids travel as integer bits."""
    import torch

    rng = np.random.default_rng(5)
    P = 50_000
    ids = (1 << 27) + 1 + rng.integers(0, 2000, P) * 2 + (rng.integers(0, 8, P) << 0) * 0  # odd node ids > 2^24
    child = rng.integers(0, 8, P)
    small = rng.random(P) < 0.3
    node = np.where(small, rng.integers(0, 1000, P), ids).astype(np.int64)
    col = np.array([mnv.lib().mnv_tracker_encode_chunk(int(v)) for v in node], np.float32)
    assert np.array_equal(mnv.decode_tracker_chunks(col), node)
    assert all(mnv.lib().mnv_tracker_decode_chunk(float(c)) == int(v) for c, v in zip(col[:200], node[:200]))
    prio = (node % 7).astype(np.float32)
    rows = np.stack([prio, col, child.astype(np.float32)], 1)
    nodes, nc = mnv.select_candidates(torch.from_numpy(rows).cuda(), 4096, "split")
    # numpy restatement on exact integers
    key = node * 8 + child
    u, c = np.unique(key, return_counts=True)
    keep = c >= 2
    u, c = u[keep], c[keep]
    pr = (u >> 3) % 7
    order = np.lexsort((u, pr, -c))
    want = np.stack([u[order] >> 3, u[order] & 7], 1)[:4096].astype(np.int32)
    assert nc == len(u) and np.array_equal(nodes.cpu().numpy(), want)
    assert (nodes.cpu().numpy()[:, 0] > (1 << 24)).any()


def test_sharded_commit_equals_fused_commit(mnv):
    """Multi-GPU refinement: children reduced in rank-sized pieces and committed from the gathered payload records
    give the tree the fused mnv_tree_commit_children gives — data, child, parent and counts bit for bit."""
    import torch

    tree = mnv.synth.make_tree(depth=5)
    opt = mnv.default_options(background_brightness=0.0, samples_per_corner=6)  # 6: mean is not an exact scale
    k, c, D = 37, 6, tree.data_dim
    rng = np.random.default_rng(9)
    dts = [mnv.DeviceTree(tree, max_capacity=tree.capacity + 64) for _ in range(2)]
    cam = mnv.synth.default_camera(160, 90, pose=1)
    ts = torch.empty((160 * 90, 3), device="cuda")
    tp = torch.empty((160 * 90, 3), device="cuda")
    dts[0].render(cam, opt, to_split=ts, to_sample=tp)
    nodes, _ = mnv.select_candidates(ts, k, "split")
    k = nodes.shape[0]
    assert k > 8
    results = torch.from_numpy(rng.standard_normal((k * 8, c, D + 1)).astype(np.float32)).cuda()
    for dt in dts:
        samples = torch.rand((k * 8, c, 3), device="cuda")
        cluster = torch.zeros((k * 8, c), dtype=torch.int16, device="cuda")
        dt.add_children(opt, nodes, samples, cluster, [1, 1], [-1.0] * 3, [2.0] * 3)
    dts[0].commit_children(opt, k, results)
    world = 3
    per = (k * 8 + world - 1) // world
    pieces = []
    for r in range(world):
        lo, hi = min(r * per, k * 8), min((r + 1) * per, k * 8)
        piece = torch.zeros((per, dts[1].record_bytes), dtype=torch.uint8, device="cuda")
        if hi > lo:
            piece[: hi - lo] = dts[1].reduce_children(opt, results[lo:hi].contiguous())
        pieces.append(piece)
    dts[1].commit_children_records(opt, k, torch.cat(pieces))
    a, b = dts[0].download(), dts[1].download()
    assert dts[0].capacity == dts[1].capacity == tree.capacity + k
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    # and the march sees the same leaves
    ia = dts[0].render_logged(cam, opt)
    ib = dts[1].render_logged(cam, opt)
    assert np.array_equal(ia["hash"], ib["hash"]) and np.array_equal(ia["rgba"], ib["rgba"])
    for dt in dts:
        dt.close()


def test_query_submodules_dispatch(mnv):
    import torch
    from mlp_reference import MegaNerfMLP

    subs = []
    refs = []
    for s in range(3):
        torch.manual_seed(10 + s)
        r = MegaNerfMLP().cuda().eval()
        refs.append(r)
        subs.append(r.export())
    model = mnv.MlpModel(subs, grid_dim=(1, 3))
    V = 5000
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand((V, model.in_dim), device="cuda", generator=g) * 2 - 1
    x[:, -1] = 0
    cluster = torch.randint(0, 3, (V,), device="cuda", generator=g).to(torch.int16)
    out = torch.full((V, model.out_dim + 1), -7.0, device="cuda")
    model.query_submodules(cluster, x, out)
    torch.cuda.synchronize()
    assert (out[:, -1] == -7.0).all()  # scatter_ writes only the model's columns
    for s in range(3):
        m = cluster == s
        with torch.no_grad():
            want = refs[s](x[m], emulate_bf16=True)
        got = out[m][:, : model.out_dim]
        rel = (got - want).norm() / want.norm()
        assert rel < 1e-3, (s, rel.item())
        # identical to running that sub-module directly on the gathered rows (same kernel, same tiles?)
        direct = model.forward(x[m].contiguous(), submodule=s)
        assert (direct - got).norm() / direct.norm() < 1e-3
    model.close()


def test_refinement_round_trip(mnv):
    """render (votes) -> select -> add children -> MLP -> commit: the tree grows and still renders."""
    import torch
    from mlp_reference import MegaNerfMLP

    tree = mnv.synth.make_tree(depth=6)
    cap = tree.capacity
    dt = mnv.DeviceTree(tree, max_capacity=cap + 5000)
    torch.manual_seed(3)
    model = mnv.MlpModel([MegaNerfMLP().export()])
    cam = mnv.synth.default_camera(320, 180, pose=1)
    opt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8], use_splitting=True,
                              appearance_embedding=0, split_batch_size=512)
    P = 320 * 180
    ts = torch.empty((P, 3), device="cuda")
    tp = torch.empty((P, 3), device="cuda")
    before = dt.render(cam, opt, to_split=ts, to_sample=tp).cpu().numpy()
    nodes, _ = mnv.select_candidates(ts, opt.split_batch_size, "split")
    n, c, rd = nodes.shape[0], opt.samples_per_corner, 4
    assert n > 0
    samples = torch.rand((n * 8, c, rd), device="cuda")
    cluster = torch.zeros((n * 8, c), dtype=torch.int16, device="cuda")
    dt.add_children(opt, nodes, samples, cluster, [1, 1], [-1, -1, -1], [2, 2, 2])
    results = torch.zeros((n * 8 * c, tree.data_dim + 1), device="cuda")
    model.query_submodules(cluster.view(-1), samples.view(-1, rd), results)
    dt.commit_children(opt, n, results.view(n * 8, c, -1))
    torch.cuda.synchronize()
    assert dt.capacity == cap + n
    after = dt.render(cam, opt).cpu().numpy()
    assert after.shape == before.shape and (after[..., 3] == 255).all()
    data, child, parent, counts = dt.download()
    assert (child[nodes.cpu().numpy()[:, 0], nodes.cpu().numpy()[:, 1]] > 0).all()
    assert np.isfinite(data[cap:].astype(np.float32)).all()
    model.close()
    dt.close()
