"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's refinement ops
(integer / index work + the sample geometry):
  add_children_and_generate_samples_kernel  src/cuda/renderer_kernel.cu:170-198
  generate_samples_inner                    src/cuda/renderer_kernel.cu:88-168
  adjust_parents_and_children_kernel        src/cuda/renderer_kernel.cu:63-86
  Impl::prune_tree gather                   src/renderer/cuda_renderer.cpp:360-377
Operates on the reference's AoS host arrays (child relative offsets, packed parent)."""
import numpy as np


def voxel_corner(parent, node, child):
    """corner of voxel (node, child) in tree space [0,1)^3 and its depth (renderer_kernel.cu:101-121)."""
    corners = np.zeros(3, np.float32)
    cur, depth = node * 8 + child, 0
    while True:
        k, j, i = cur & 1, (cur >> 1) & 1, (cur >> 2) & 1
        nd = cur >> 3
        corners = ((corners + np.array([i, j, k], np.float32)) / np.float32(2)).astype(np.float32)
        if nd == 0:
            break
        cur = int(parent[nd])
        depth += 1
    return corners, depth


def generate_samples(parent, scale, offset, packed_voxels, rand, need_viewdir, appearance, grid_dim,
                     min_position, rng):
    """rand f32 [m, c, rand_dim] in [0,1) -> (samples, cluster i16 [m, c])."""
    out = rand.astype(np.float32).copy()
    m, c, _ = out.shape
    cluster = np.zeros((m, c), np.int16)
    for v in range(m):
        node, child = divmod(int(packed_voxels[v]), 8)
        corners, depth = voxel_corner(parent, node, child)
        length = np.float32(2.0 ** (-depth - 1))
        for a in range(3):
            corner = np.float32((corners[a] - np.float32(offset[a])) / np.float32(scale[a]))
            k = np.float32(length / np.float32(scale[a]))
            # FFMA: single rounding of s*k + corner
            out[v, :, a] = (out[v, :, a].astype(np.float64) * np.float64(k) + np.float64(corner)).astype(np.float32)
        col = 3
        if need_viewdir:
            out[v, :, 3:6] = [1, 0, 0]
            col = 6
        if appearance != -1:
            out[v, :, col] = appearance
        g0, g1 = np.float32(grid_dim[0]), np.float32(grid_dim[1])
        a_ = np.maximum(np.minimum((out[v, :, 1] - np.float32(min_position[1])) / np.float32(rng[1]) * g0, g0 - 1), 0)
        b_ = np.maximum(np.minimum((out[v, :, 2] - np.float32(min_position[2])) / np.float32(rng[2]) * g1, g1 - 1), 0)
        cluster[v] = a_.astype(np.int32) * int(grid_dim[1]) + b_.astype(np.int32)
    return out, cluster


def add_children(child, parent, capacity, parent_nodes):
    """-> (child', parent') with n new nodes appended at capacity.. (links only)."""
    n = len(parent_nodes)
    child2 = np.concatenate([child[:capacity], np.zeros((n, 8), np.int32)])
    parent2 = np.concatenate([parent[:capacity], np.zeros(n, np.int32)])
    for r, (pn, pc) in enumerate(parent_nodes):
        a = capacity + r
        child2[pn, pc] = a - pn
        parent2[a] = pn * 8 + pc
    return child2, parent2


def prune(child, parent, data, to_delete):
    """Impl::prune_tree on host arrays -> (child', parent', data', kept_index)."""
    to_delete = np.asarray(to_delete, bool)
    cap = len(to_delete)
    shifts = np.cumsum(to_delete).astype(np.int32)
    child, parent = child[:cap].copy(), parent[:cap].copy()
    first = int(np.argmin(shifts))
    for chunk in range(first, cap):
        pn, pc = divmod(int(parent[chunk]), 8)
        if chunk == 0:
            continue
        if to_delete[chunk]:
            child[pn, pc] = 0
        else:
            child[pn, pc] += shifts[pn] - shifts[chunk]
            parent[chunk] -= shifts[pn] * 8
    keep = np.nonzero(~to_delete)[0]
    return child[keep], parent[keep], data[:cap][keep], keep
