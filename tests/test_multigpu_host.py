"""CPU tests (gloo, world_size 2 and 4) of the host side of the multi-GPU paths: pixel-block partition,
cell boxes, tree restriction and the handle exchange of mega_nerf_viewer_b200.multigpu."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_owner_ranges_tile_the_frame(mnv):
    MG = mnv.multigpu
    for P in (1, 7, 1000, 1920 * 1080, 3840 * 2160 + 3):
        for world in (1, 2, 4, 8):
            ranges = [MG.owner_range(P, world, r) for r in range(world)]
            assert sum(n for _, n in ranges) == P
            b = MG.block_pixels(P, world)
            for r, (first, n) in enumerate(ranges):
                assert first == min(r * b, P) and 0 <= n <= b


def test_cell_boxes_follow_the_cluster_rule(mnv):
    """cluster = floor((y - min_y) / range_y * g0) * g1 + floor((z - min_z) / range_z * g1) (rt_core.cuh:541-549)."""
    S = mnv.synth
    for world in (1, 2, 4, 8):
        g = S.grid_for_world(world)
        boxes = S.cell_boxes(g, world)
        assert boxes.shape == (world, 6) and g[0] * g[1] == world
        rng = np.random.default_rng(world)
        pts = rng.random((2000, 3))
        cid = np.minimum((pts[:, 1] * g[0]).astype(int), g[0] - 1) * g[1] + np.minimum((pts[:, 2] * g[1]).astype(int), g[1] - 1)
        inside = np.all((pts[:, None, :] >= boxes[None, :, :3]) & (pts[:, None, :] < boxes[None, :, 3:]), -1)
        assert (inside.sum(1) == 1).all() and (inside.argmax(1) == cid).all()


def test_restricted_trees_partition_the_leaves(mnv):
    S = mnv.synth
    tree = S.make_tree(depth=6)
    boxes = S.cell_boxes((2, 4))
    shaded_full = int((tree.data[..., -1].astype(np.float32) > 0).sum())
    shaded = 0
    for b in boxes:
        sub = S.restrict_tree(tree, b)
        assert sub.capacity < tree.capacity
        assert (sub.child >= 0).all() and sub.parent[0] == 0
        shaded += int((sub.data[..., -1].astype(np.float32) > 0).sum())
    assert shaded == shaded_full  # every occupied leaf lives in exactly one cell's tree


@pytest.mark.parametrize("world", [2, 4])
def test_gloo_workers(world):
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29540 + world),
                        os.path.join(ROOT, "tests", "multigpu_worker.py"), "--backend", "gloo", "--depth", "5"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert j["world"] == world and j["ok"] is True
    assert sum(j["nodes"]) < j["full_nodes"] * 1.5
