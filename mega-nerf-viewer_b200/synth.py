"""Synthetic N3Tree generator.

Writes / returns trees in the reference's on-disk schema — the keys
``N3Tree::load_npz`` reads (src/n3tree/n3tree.cpp:28-205):

  data_dim      i64 scalar
  data_format   '<U*' string ("SH9", "RGBA", ...)
  invradius3    f32[3]
  offset        f32[3]
  child         i32[cap,2,2,2]   relative offset to the child node, 0 = leaf
  parent_depth  i32[cap,2]       (packed parent slot node*8+child, depth)
  data          f16[cap,2,2,2,data_dim]

Scene: world box [-1,1]^3 (invradius3 = offset = 0.5).  A cell is refined while
its centre lies within 1.5 cell-diagonals (vertical distance) of a seeded
height field z = f(x, y) (sum of 6 sinusoids, seed 0) until ``depth``.  SH
coefficients ~ N(0,1) (seed 1).  sigma = 0 in non-surface leaves and U(5,50)
in finest-level leaves near the surface.

Nodes are laid out breadth-first, so every relative child offset is positive.
There are no test fixtures in the reference; this generator is
the workload definition for parity tests and for bench.py ("data": "synthetic").
"""
from __future__ import annotations

import dataclasses
import numpy as np


@dataclasses.dataclass
class HostTree:
    """Host arrays in the reference's AoS layout (n3tree.cpp:82-107,177-193)."""

    N: int
    data_dim: int
    data_format: str
    child: np.ndarray  # i32 [cap, 8]
    parent: np.ndarray  # i32 [cap]  packed parent slot (node*8 + child)
    depth: np.ndarray  # i32 [cap]   depth column of parent_depth (unused by the viewer)
    data: np.ndarray  # f16 [cap, 8, data_dim]
    scale: np.ndarray  # f32 [3] (invradius3)
    offset: np.ndarray  # f32 [3]

    @property
    def capacity(self) -> int:
        return int(self.child.shape[0])

    @property
    def basis_dim(self) -> int:
        fmt = self.data_format
        digits = "".join(ch for ch in fmt if ch.isdigit())
        return int(digits) if digits else -1

    def nbytes(self) -> int:
        return self.child.nbytes + self.parent.nbytes + self.data.nbytes

    def save_npz(self, path: str, compressed: bool = False) -> None:
        cap = self.capacity
        arrays = dict(
            data_dim=np.int64(self.data_dim),
            data_format=np.array(self.data_format),
            invradius3=self.scale.astype(np.float32),
            offset=self.offset.astype(np.float32),
            child=self.child.reshape(cap, 2, 2, 2).astype(np.int32),
            parent_depth=np.stack([self.parent, self.depth], axis=1).astype(np.int32),
            data=self.data.reshape(cap, 2, 2, 2, self.data_dim).astype(np.float16),
        )
        (np.savez_compressed if compressed else np.savez)(path, **arrays)

    @staticmethod
    def load_npz(path: str) -> "HostTree":
        """numpy-side reader used only by tests to cross-check the C++ loader."""
        z = np.load(path)
        cap = z["child"].shape[0]
        data_dim = int(z["data_dim"])
        if "invradius3" in z:
            scale = z["invradius3"].astype(np.float32)
        else:
            scale = np.full(3, np.float32(z["invradius"]), np.float32)
        pd = z["parent_depth"]
        return HostTree(
            N=int(z["child"].shape[1]),
            data_dim=data_dim,
            data_format=str(z["data_format"]),
            child=np.ascontiguousarray(z["child"].reshape(cap, 8).astype(np.int32)),
            parent=np.ascontiguousarray(pd[:, 0].astype(np.int32)),
            depth=np.ascontiguousarray(pd[:, 1].astype(np.int32)),
            data=np.ascontiguousarray(z["data"].reshape(cap, 8, data_dim)),
            scale=scale,
            offset=z["offset"].astype(np.float32),
        )


def height_field(x: np.ndarray, y: np.ndarray, seed: int = 0, tilt: float = 0.0, amp_scale: float = 1.0) -> np.ndarray:
    """z = f(x, y): sum of 6 seeded sinusoids, |z| <= ~0.55 in world units (times amp_scale), plus tilt * x.
    With tilt ~0.75 and amp_scale ~0.4 the surface crosses the whole z extent, so the cells of a (y, z)
    sub-module grid hold equal shares of it (the flat default leaves the outer z cells empty)."""
    rng = np.random.default_rng(seed)
    amp = rng.uniform(0.04, 0.14, 6)
    kx = rng.uniform(-7.0, 7.0, 6)
    ky = rng.uniform(-7.0, 7.0, 6)
    ph = rng.uniform(0.0, 2 * np.pi, 6)
    z = np.zeros_like(x, dtype=np.float64)
    for a, u, v, p in zip(amp, kx, ky, ph):
        z += a * np.sin(u * x + v * y + p)
    if amp_scale != 1.0:
        z *= amp_scale
    if tilt != 0.0:
        z += tilt * x
    return z


def _near_surface(ix, iy, iz, level, seed, margin=1.5, tilt=0.0, amp_scale=1.0):
    """Cells (integer coords at `level`) whose centre is within `margin`
    cell-diagonals (vertical distance) of the height field. World box [-1,1]^3."""
    size = 2.0 / (1 << level)
    cx = (ix + 0.5) * size - 1.0
    cy = (iy + 0.5) * size - 1.0
    cz = (iz + 0.5) * size - 1.0
    diag = np.sqrt(3.0) * size
    return np.abs(cz - height_field(cx, cy, seed, tilt, amp_scale)) < margin * diag


def make_tree(
    depth: int = 8,
    data_format: str = "SH9",
    seed: int = 0,
    blocks_yz: tuple[int, int] | None = None,
    block_depths: list[int] | None = None,
    sigma_range: tuple[float, float] = (5.0, 50.0),
    max_nodes: int | None = None,
    tilt: float = 0.0,
    amp_scale: float = 1.0,
    fast_data: bool = False,
) -> HostTree:
    """Build the synthetic octree.

    depth         finest leaf depth in the reference's counting (root's children
                  are depth 1), i.e. internal nodes exist on levels 0..depth-1.
    blocks_yz     config 3: (gy, gz) grid of Mega-NeRF spatial blocks on the
                  (y, z) axes; block b may use its own ``block_depths[b]``.
    tilt, amp_scale  shape of the height field (see height_field).
    fast_data     multi-GB trees: the N(0,1) coefficients repeat a 2^16-node random block (rolled per
                  chunk) instead of drawing ~4*10^9 normals; sigma is still per leaf.  Traversal does not
                  depend on the coefficients.
    """
    basis = "".join(ch for ch in data_format if ch.isdigit())
    is_sh = data_format.upper().startswith("SH")
    data_dim = 3 * int(basis) + 1 if is_sh else 4

    # level-by-level BFS over integer cell coordinates
    lv_coords = [(np.zeros(1, np.int64), np.zeros(1, np.int64), np.zeros(1, np.int64))]
    lv_child_refined = []  # per level: bool [n_l, 8]
    lv_surface = []  # per level: bool [n_l, 8] finest-level near-surface leaves
    oi, oj, ok = np.meshgrid([0, 1], [0, 1], [0, 1], indexing="ij")
    oi, oj, ok = oi.ravel(), oj.ravel(), ok.ravel()  # child index = 4i + 2j + k (i <-> x)
    total = 1
    for level in range(depth):
        ix, iy, iz = lv_coords[level]
        cx = (ix[:, None] * 2 + oi[None, :])
        cy = (iy[:, None] * 2 + oj[None, :])
        cz = (iz[:, None] * 2 + ok[None, :])
        near = _near_surface(cx, cy, cz, level + 1, seed, tilt=tilt, amp_scale=amp_scale)
        lim = depth
        if blocks_yz is not None and block_depths is not None:
            gy, gz = blocks_yz
            n = 1 << (level + 1)
            by = np.minimum(cy * gy // n, gy - 1)
            bz = np.minimum(cz * gz // n, gz - 1)
            lim = np.asarray(block_depths, np.int64)[by * gz + bz]
        refine = near & ((level + 1) < lim)
        if max_nodes is not None and total + int(refine.sum()) > max_nodes:
            refine = np.zeros_like(refine)
        surface = near & ~refine & ((level + 1) == lim)
        lv_child_refined.append(refine)
        lv_surface.append(surface)
        sel = np.nonzero(refine.ravel())[0]
        total += sel.size
        lv_coords.append((cx.ravel()[sel], cy.ravel()[sel], cz.ravel()[sel]))
        if sel.size == 0:
            break

    n_levels = len(lv_child_refined)
    counts = [lv_coords[l][0].size for l in range(n_levels)]
    starts = np.concatenate([[0], np.cumsum(counts)])
    cap = int(starts[n_levels])
    child = np.zeros((cap, 8), np.int32)
    parent = np.zeros(cap, np.int32)
    depth_col = np.zeros(cap, np.int32)
    surface_mask = np.zeros((cap, 8), bool)
    for level in range(n_levels):
        s, e = int(starts[level]), int(starts[level + 1])
        refine = lv_child_refined[level]
        surface_mask[s:e] = lv_surface[level]
        depth_col[s:e] = level
        flat = np.nonzero(refine.ravel())[0]
        if flat.size == 0 or level + 1 >= n_levels:
            continue
        next_ids = int(starts[level + 1]) + np.arange(flat.size, dtype=np.int64)
        node = s + flat // 8
        slot = flat % 8
        child[node, slot] = (next_ids - node).astype(np.int32)
        parent[next_ids] = (node * 8 + slot).astype(np.int32)
    parent[0] = 0

    rng = np.random.default_rng(seed + 1)
    data = np.empty((cap, 8, data_dim), np.float16)
    step = 1 << (16 if fast_data else 18)
    base = None
    for s in range(0, cap, step):
        e = min(cap, s + step)
        if fast_data:
            if base is None:
                base = rng.standard_normal((step, 8, data_dim), dtype=np.float32)
            blk = np.roll(base, (s // step) * 977, axis=0)[: e - s].copy()
        else:
            blk = rng.standard_normal((e - s, 8, data_dim), dtype=np.float32)
        if not is_sh:
            blk[..., :3] = 1.0 / (1.0 + np.exp(-blk[..., :3]))  # RGBA stores colours directly
        sig = rng.uniform(sigma_range[0], sigma_range[1], (e - s, 8)).astype(np.float32)
        blk[..., data_dim - 1] = np.where(surface_mask[s:e], sig, 0.0)
        data[s:e] = blk.astype(np.float16)

    return HostTree(
        N=2,
        data_dim=data_dim,
        data_format=data_format,
        child=child,
        parent=parent,
        depth=depth_col,
        data=data,
        scale=np.full(3, 0.5, np.float32),
        offset=np.full(3, 0.5, np.float32),
    )


def brute_force_query(tree: HostTree, xyz: np.ndarray):
    """Integer-coordinate descent, independent of the float recurrence in
    query_single_from_root (include/cuda/rt_core.cuh:117-159): returns
    (chunk, child, depth) per point for tree-space xyz in [0,1)."""
    xyz = np.clip(xyz.astype(np.float32), np.float32(0.0), np.float32(1.0) - np.float32(1e-6))
    q = np.floor(xyz.astype(np.float64) * (1 << 24)).astype(np.int64)  # exact: xyz is fp32 < 1
    n = xyz.shape[0]
    node = np.zeros(n, np.int64)
    out = np.zeros((n, 3), np.int32)
    active = np.ones(n, bool)
    level = 1
    while active.any():
        sh = 24 - level
        bits = (q >> sh) & 1
        cidx = bits[:, 0] * 4 + bits[:, 1] * 2 + bits[:, 2]
        skip = tree.child[node, cidx]
        leaf = active & (skip == 0)
        out[leaf, 0] = node[leaf]
        out[leaf, 1] = cidx[leaf]
        out[leaf, 2] = level
        active &= skip != 0
        node = np.where(active, node + skip, node)
        level += 1
        if level > 24:
            raise RuntimeError("tree deeper than 24 levels")
    return out


def default_camera(width: int = 1920, height: int = 1080, pose: int = 0, n_poses: int = 16):
    """Config-2 camera: viewer defaults (main.cpp:491-504,
    src/camera.cpp:41-44) scaled to the frame; `pose` steps a 16-pose orbit about z.
    Returns dict(width,height,fx,fy,cx,cy,c2w[12]) with c2w = right,up,back,center."""
    ang = 2.0 * np.pi * pose / n_poses
    c, s = np.cos(ang), np.sin(ang)
    rot = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    center = rot @ (np.array([-3.5, 0.0, 3.5]) * 0.5)
    back = rot @ np.array([-0.7071068, 0.0, 0.7071068])
    world_up = np.array([0.0, 0.0, 1.0])
    # Camera::_update, src/camera.cpp:54-82 (float32 like glm)
    back = (back / np.linalg.norm(back)).astype(np.float32)
    right = np.cross(world_up, back)
    right = (right / np.linalg.norm(right)).astype(np.float32)
    up = np.cross(back, right).astype(np.float32)
    fx = 1111.0 * (width / 800.0)
    return dict(
        width=width, height=height, fx=float(fx), fy=float(fx), cx=width / 2.0, cy=height / 2.0,
        c2w=np.concatenate([right, up, back, center.astype(np.float32)]).astype(np.float32),
    )


def make_mlp_weights(seed: int = 3, basis_dim: int = 9, n_appearance: int = 4, appearance_dim: int = 48,
                     need_viewdir: bool = False, width: int = 256, n_layers: int = 8, skip_layer: int = 4,
                     pe_xyz: int = 12, pe_dir: int = 4, head_width: int = 128, sigma_activation: int = 1) -> dict:
    """Random-init Mega-NeRF sub-MLP of the shapes BASELINE.json names,
    PyTorch's default initialisers restated in numpy (nn.Linear: U(-1/sqrt(in), 1/sqrt(in)) for weight
    and bias; nn.Embedding: N(0, 1)).  Returns the dict mega_nerf_viewer_b200.MlpModel takes."""
    rng = np.random.default_rng(seed)

    def linear(n_out, n_in):
        k = 1.0 / np.sqrt(n_in)
        return (rng.uniform(-k, k, (n_out, n_in)).astype(np.float32), rng.uniform(-k, k, n_out).astype(np.float32))

    pe = 3 + 6 * pe_xyz
    trunk = [linear(width, pe if i == 0 else (width + pe if i == skip_layer else width)) for i in range(n_layers)]
    sigma_w, sigma_b = linear(1, width)
    final_w, final_b = linear(width, width)
    head_in = width + ((3 + 6 * pe_dir) if need_viewdir else 0) + appearance_dim
    head1_w, head1_b = linear(head_width, head_in)
    head2_w, head2_b = linear(3 * basis_dim, head_width)
    return dict(
        trunk_w=[w for w, _ in trunk], trunk_b=[b for _, b in trunk], sigma_w=sigma_w, sigma_b=sigma_b,
        final_w=final_w, final_b=final_b,
        embedding=rng.standard_normal((n_appearance, appearance_dim)).astype(np.float32) if appearance_dim > 0 else None,
        head1_w=head1_w, head1_b=head1_b, head2_w=head2_w, head2_b=head2_b, skip_layer=skip_layer,
        pe_xyz_freqs=pe_xyz, pe_dir_freqs=pe_dir, need_viewdir=need_viewdir, sigma_activation=sigma_activation)


def cell_boxes(grid_dim, n_cells: int | None = None) -> np.ndarray:
    """Tree-space boxes [n, 6] = (lo xyz, hi xyz) of the Mega-NeRF (y, z) cluster grid (the rule of
    rt_core.cuh:541-549: cell = gy * grid_dim[1] + gz over the y and z extents; x is not split)."""
    g0, g1 = int(grid_dim[0]), int(grid_dim[1])
    out = []
    for gy in range(g0):
        for gz in range(g1):
            out.append([0.0, gy / g0, gz / g1, 1.0, (gy + 1) / g0, (gz + 1) / g1])
    out = np.asarray(out, np.float32)
    return out if n_cells is None else out[:n_cells]


def grid_for_world(world: int):
    """(y, z) cell grid used when the sub-modules are sharded over `world` GPUs."""
    return {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}[world]


def restrict_tree(tree: HostTree, box) -> HostTree:
    """The part of `tree` a GPU that owns the tree-space `box` needs: every subtree that does not
    overlap the box is replaced by one empty leaf (sigma 0).  Leaf data inside the box, and the
    leaf-visit sequence of any ray clipped to the box, are unchanged."""
    lo, hi = np.asarray(box[:3], np.float64), np.asarray(box[3:], np.float64)
    D = tree.data_dim
    nodes = np.array([0], np.int64)         # old node ids of the current level
    coords = np.zeros((1, 3), np.int64)     # cell coordinates of those nodes at their level
    level = 0
    new_child, new_parent, new_depth, new_data = [], [], [], []
    n_new = 1
    new_id_of = [np.array([0], np.int64)]
    parents_packed = [np.array([0], np.int64)]
    while nodes.size:
        bits = np.stack(np.meshgrid([0, 1], [0, 1], [0, 1], indexing="ij"), -1).reshape(8, 3)
        ccoord = coords[:, None, :] * 2 + bits[None]                    # [n, 8, 3] at level+1
        size = 1.0 / (1 << (level + 1))
        clo, chi = ccoord * size, (ccoord + 1) * size
        overlap = np.all((clo < hi) & (chi > lo), axis=-1)                # [n, 8]
        rel = tree.child[nodes]                                           # [n, 8]
        keep = (rel != 0) & overlap
        data = tree.data[nodes].copy()
        cut = (rel != 0) & ~overlap
        data[cut] = 0                                                     # pruned subtree -> empty leaf
        my_ids = new_id_of[-1]
        kn, kc = np.nonzero(keep)
        next_ids = n_new + np.arange(kn.size, dtype=np.int64)
        ch = np.zeros_like(rel)
        ch[kn, kc] = (next_ids - my_ids[kn]).astype(np.int32)
        new_child.append(ch)
        new_data.append(data)
        new_depth.append(np.full(nodes.size, level, np.int32))
        new_parent.append(parents_packed[-1])
        parents_packed.append(my_ids[kn] * 8 + kc)
        new_id_of.append(next_ids)
        n_new += kn.size
        nodes = nodes[kn] + rel[kn, kc]
        coords = ccoord[kn, kc]
        level += 1
    return HostTree(N=2, data_dim=D, data_format=tree.data_format,
                    child=np.concatenate(new_child).astype(np.int32),
                    parent=np.concatenate(new_parent).astype(np.int32),
                    depth=np.concatenate(new_depth).astype(np.int32),
                    data=np.concatenate(new_data).astype(np.float16),
                    scale=tree.scale.copy(), offset=tree.offset.copy())
