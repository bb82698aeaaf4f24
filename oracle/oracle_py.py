"""TEST INFRASTRUCTURE ONLY — ctypes loader for oracle/liboracle.so (the CPU
restatement) and oracle/_ref/libref_render*.so (the reference's own CUDA path).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_render.so")
REF_INSTR_SO = os.path.join(HERE, "_ref", "libref_render_instr.so")


class RenderOptions(C.Structure):
    """include/mnv_b200.h mnv_render_options == include/render_options.hpp:9-56."""

    _fields_ = [
        ("step_size", C.c_float),
        ("sigma_thresh", C.c_float),
        ("stop_thresh", C.c_float),
        ("background_brightness", C.c_float),
        ("render_bbox", C.c_float * 6),
        ("basis_minmax", C.c_int * 2),
        ("rot_dirs", C.c_float * 3),
        ("show_grid", C.c_bool),
        ("grid_max_depth", C.c_int),
        ("render_depth", C.c_bool),
        ("use_splitting", C.c_bool),
        ("use_guided_sampling", C.c_bool),
        ("max_depth", C.c_int),
        ("samples_per_corner", C.c_int),
        ("split_batch_size", C.c_int),
        ("nerf_batch_size", C.c_int),
        ("max_sample_count", C.c_int),
        ("need_viewdir", C.c_bool),
        ("appearance_embedding", C.c_int),
        ("max_guided_samples", C.c_int),
    ]


def default_options(**kw) -> RenderOptions:
    """Struct defaults of include/render_options.hpp:12-55."""
    o = RenderOptions()
    o.step_size = 1e-4
    o.sigma_thresh = 1e-2
    o.stop_thresh = 1e-2
    o.background_brightness = 1.0
    o.render_bbox[:] = [0, 0, 0, 1, 1, 1]
    o.basis_minmax[:] = [0, 24]
    o.rot_dirs[:] = [0, 0, 0]
    o.show_grid = False
    o.grid_max_depth = 4
    o.render_depth = False
    o.use_splitting = False
    o.use_guided_sampling = False
    o.max_depth = 16
    o.samples_per_corner = 8
    o.split_batch_size = 4192
    o.nerf_batch_size = 1024
    o.max_sample_count = 256
    o.need_viewdir = False
    o.appearance_embedding = -1
    o.max_guided_samples = 128
    for k, v in kw.items():
        if isinstance(v, (list, tuple, np.ndarray)):
            getattr(o, k)[:] = list(v)
        else:
            setattr(o, k, v)
    return o


class Camera(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("c2w", C.c_float * 12),
    ]


def make_camera(d: dict) -> Camera:
    c = Camera()
    c.width, c.height = int(d["width"]), int(d["height"])
    c.fx, c.fy, c.cx, c.cy = d["fx"], d["fy"], d["cx"], d["cy"]
    c.c2w[:] = [float(v) for v in d["c2w"]]
    return c


class TreeDesc(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("data_dim", C.c_int), ("format", C.c_int), ("basis_dim", C.c_int),
        ("capacity", C.c_int64),
        ("data", C.c_void_p), ("child", C.c_void_p), ("parent", C.c_void_p),
        ("sample_counts", C.c_void_p),
        ("scale", C.c_float * 3), ("offset", C.c_float * 3),
    ]


def make_tree_desc(tree, sample_counts: np.ndarray | None = None):
    """tree: synth.HostTree.  Returns (desc, keepalive)."""
    d = TreeDesc()
    d.N = tree.N
    d.data_dim = tree.data_dim
    d.format = 1 if tree.data_format.upper().startswith("SH") else 0
    d.basis_dim = tree.basis_dim
    d.capacity = tree.capacity
    data = np.ascontiguousarray(tree.data.view(np.uint16))
    child = np.ascontiguousarray(tree.child, np.int32)
    parent = np.ascontiguousarray(tree.parent, np.int32)
    d.data = data.ctypes.data
    d.child = child.ctypes.data
    d.parent = parent.ctypes.data
    keep = [data, child, parent]
    if sample_counts is not None:
        sc = np.ascontiguousarray(sample_counts, np.int16)
        d.sample_counts = sc.ctypes.data
        keep.append(sc)
    d.scale[:] = [float(v) for v in tree.scale]
    d.offset[:] = [float(v) for v in tree.offset]
    return d, keep


def build_oracle() -> None:
    subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, capture_output=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        _lib = C.CDLL(ORACLE_SO)
        _lib.oracle_query_points.restype = C.c_int
        _lib.oracle_query_points.argtypes = [C.POINTER(TreeDesc), C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        _lib.oracle_render_voxels.restype = C.c_int
        _lib.oracle_render_voxels.argtypes = [
            C.POINTER(TreeDesc), C.POINTER(Camera), C.POINTER(RenderOptions),
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
            C.c_int, C.c_int, C.c_int, C.c_int,
        ]
        _lib.oracle_num_threads.restype = C.c_int
        _lib.oracle_get_samples.restype = C.c_int
        _lib.oracle_get_samples.argtypes = [C.POINTER(TreeDesc), C.POINTER(Camera), C.POINTER(RenderOptions),
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        _lib.oracle_composite_nerf.restype = C.c_int
        _lib.oracle_composite_nerf.argtypes = [C.c_int, C.c_int, C.POINTER(Camera), C.POINTER(RenderOptions),
                                               C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data


def query_points(tree, xyz: np.ndarray, nthreads: int = 0) -> np.ndarray:
    desc, keep = make_tree_desc(tree)
    xyz = np.ascontiguousarray(xyz, np.float32)
    out = np.empty((xyz.shape[0], 3), np.int32)
    rc = lib().oracle_query_points(C.byref(desc), xyz.ctypes.data, xyz.shape[0], out.ctypes.data, nthreads)
    assert rc == 0, rc
    return out


def render_voxels(tree, cam: dict, opt: RenderOptions, *, trackers=False, log_cap=0, stats=True,
                  y0=0, y1=None, row_step=1, nthreads=0, sample_counts=None):
    """Returns dict(rgba[H,W,4], to_split, to_sample, hash, count, shaded, log)."""
    desc, keep = make_tree_desc(tree, sample_counts)
    c = make_camera(cam)
    H, W = c.height, c.width
    P = H * W
    rgba = np.zeros((H, W, 4), np.uint8)
    split = np.full((P, 3), -1, np.float32) if trackers else None
    sample = np.full((P, 3), -1, np.float32) if trackers else None
    vh = np.zeros(P, np.uint64) if stats else None
    vc = np.zeros(P, np.int32) if stats else None
    vs = np.zeros(P, np.int32) if stats else None
    vlog = np.full((P, log_cap), -1, np.int32) if log_cap > 0 else None
    rc = lib().oracle_render_voxels(
        C.byref(desc), C.byref(c), C.byref(opt), rgba.ctypes.data, None, _ptr(split), _ptr(sample),
        None, 0, 1, _ptr(vh), _ptr(vc), _ptr(vs), _ptr(vlog), log_cap,
        y0, H if y1 is None else y1, row_step, nthreads)
    assert rc == 0, rc
    return dict(rgba=rgba, to_split=split, to_sample=sample, hash=vh, count=vc, shaded=vs, log=vlog)


def sample_dim(opt: RenderOptions) -> int:
    """cuda_renderer.cpp:478-489: z + xyz (+ view dir) (+ appearance index)."""
    return 4 + (3 if opt.need_viewdir else 0) + (1 if opt.appearance_embedding != -1 else 0)


def compact_samples(num_samples, samples, cluster):
    """The reference's mask compaction + cumsum (cuda_renderer.cpp:116-120) on host arrays:
    -> offsets i64 [P], z_vals [V], rows [V, sd-1], cluster [V]."""
    P, S, sd = samples.shape
    flat = samples.reshape(-1, sd)
    valid = flat[:, 0] >= 0
    offsets = np.cumsum(num_samples.astype(np.int64))
    return offsets, flat[valid, 0].copy(), flat[valid, 1:].copy(), cluster.reshape(-1)[valid].copy()


def get_samples(tree, cam: dict, opt: RenderOptions, grid_dim, min_position, rng, nthreads=0):
    """CPU oracle of get_samples_from_voxels with the reference's dense outputs."""
    desc, keep = make_tree_desc(tree)
    c = make_camera(cam)
    P, S, sd = c.width * c.height, opt.max_guided_samples, sample_dim(opt)
    ns = np.zeros(P, np.int16)
    samples = np.full((P, S, sd), -1, np.float32)
    cluster = np.zeros((P, S), np.int16)
    split = np.full((P, 3), -1, np.float32)
    samp = np.full((P, 3), -1, np.float32)
    gd = np.ascontiguousarray(grid_dim, np.int32)
    mp = np.ascontiguousarray(min_position, np.float32)
    rg = np.ascontiguousarray(rng, np.float32)
    rc = lib().oracle_get_samples(C.byref(desc), C.byref(c), C.byref(opt), gd.ctypes.data, mp.ctypes.data,
                                  rg.ctypes.data, ns.ctypes.data, samples.ctypes.data, cluster.ctypes.data,
                                  S, sd, split.ctypes.data, samp.ctypes.data, nthreads)
    assert rc == 0, rc
    return dict(num_samples=ns, samples=samples, cluster=cluster, to_split=split, to_sample=samp)


def composite_nerf(tree, cam: dict, opt: RenderOptions, values, z_vals, offsets, sigma_col=3):
    c = make_camera(cam)
    values = np.ascontiguousarray(values, np.float32)
    z_vals = np.ascontiguousarray(z_vals, np.float32)
    offsets = np.ascontiguousarray(offsets, np.int64)
    rgba = np.zeros((c.height, c.width, 4), np.uint8)
    fmt = 1 if tree.data_format.upper().startswith("SH") else 0
    rc = lib().oracle_composite_nerf(fmt, tree.basis_dim, C.byref(c), C.byref(opt), values.ctypes.data,
                                     values.shape[1], sigma_col, z_vals.ctypes.data, offsets.ctypes.data,
                                     rgba.ctypes.data)
    assert rc == 0, rc
    return rgba


# ---------------------------------------------------------------------------
# The reference's own CUDA path (oracle/_ref, built by oracle/build_ref.sh)
# ---------------------------------------------------------------------------
def ref_available(instr: bool = False) -> bool:
    return os.path.exists(REF_INSTR_SO if instr else REF_SO)


def ref_load_host(npz_path: str) -> dict:
    """The reference's own loader (N3Tree::open + the vendored cnpy, src/n3tree/n3tree.cpp:16-205) on the host: no
    GPU needed.  Returns the arrays it built, in its layout."""
    import torch  # noqa: F401  (loads libtorch / libc10 before the driver .so)

    L = C.CDLL(REF_SO)
    L.ref_open_host.restype = C.c_void_p
    L.ref_open_host.argtypes = [C.c_char_p]
    for fn in (L.ref_close, L.ref_capacity, L.ref_data_dim):
        fn.argtypes = [C.c_void_p]
    L.ref_download.argtypes = [C.c_void_p] * 6
    L.ref_sample_counts.argtypes = [C.c_void_p, C.c_void_p]
    L.ref_format.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    sys.stdout.flush()
    saved = os.dup(1)  # the reference's loader prints its header dump to stdout
    os.dup2(2, 1)
    try:
        h = L.ref_open_host(npz_path.encode())
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    if not h:
        raise RuntimeError(f"reference N3Tree::open failed for {npz_path}")
    cap, D = L.ref_capacity(h), L.ref_data_dim(h)
    out = dict(capacity=cap, data_dim=D, data=np.empty((cap, 8, D), np.uint16), child=np.empty((cap, 8), np.int32),
               parent=np.empty(cap, np.int32), scale=np.empty(3, np.float32), offset=np.empty(3, np.float32),
               sample_counts=np.empty((cap, 8), np.int16))
    L.ref_download(h, out["data"].ctypes.data, out["child"].ctypes.data, out["parent"].ctypes.data,
                   out["scale"].ctypes.data, out["offset"].ctypes.data)
    L.ref_sample_counts(h, out["sample_counts"].ctypes.data)
    buf = C.create_string_buffer(32)
    out["basis_dim"] = L.ref_format(h, buf, 32)
    out["format"] = buf.value.decode()
    L.ref_close(h)
    return out


class RefRenderer:
    """Runs the unmodified reference kernel (rebuilt for sm_100) headlessly."""

    def __init__(self, npz_path: str, max_capacity: int = 0, instr: bool = False):
        import torch  # noqa: F401  (loads libtorch / libc10 before the driver .so)

        so = REF_INSTR_SO if instr else REF_SO
        if not os.path.exists(so):
            raise FileNotFoundError(so)
        self.instr = instr
        self.L = C.CDLL(so)
        L = self.L
        L.ref_open.restype = C.c_void_p
        L.ref_open.argtypes = [C.c_char_p, C.c_long]
        L.ref_close.argtypes = [C.c_void_p]
        L.ref_capacity.argtypes = [C.c_void_p]
        L.ref_data_dim.argtypes = [C.c_void_p]
        L.ref_download.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.ref_render_voxels.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_void_p]
        L.ref_render_frame_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_int, C.c_void_p]
        if instr:
            L.ref_render_voxels_logged.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                   C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                   C.c_void_p, C.c_void_p, C.c_int]
        self.h = L.ref_open(npz_path.encode(), max_capacity)
        if not self.h:
            raise RuntimeError(f"reference N3Tree::open failed for {npz_path}")

    def close(self):
        if getattr(self, "h", None):
            self.L.ref_close(self.h)
            self.h = None

    __del__ = close

    @property
    def capacity(self):
        return self.L.ref_capacity(self.h)

    def download(self):
        cap, D = self.capacity, self.L.ref_data_dim(self.h)
        data = np.empty((cap, 8, D), np.uint16)
        child = np.empty((cap, 8), np.int32)
        parent = np.empty(cap, np.int32)
        scale = np.empty(3, np.float32)
        offset = np.empty(3, np.float32)
        self.L.ref_download(self.h, data.ctypes.data, child.ctypes.data, parent.ctypes.data,
                            scale.ctypes.data, offset.ctypes.data)
        return data.view(np.float16), child, parent, scale, offset

    @staticmethod
    def _cam_args(cam: dict):
        intr = np.array([cam["fx"], cam["fy"], cam["cx"], cam["cy"]], np.float32)
        c2w = np.ascontiguousarray(cam["c2w"], np.float32)
        return int(cam["width"]), int(cam["height"]), intr, c2w

    def render(self, cam: dict, opt: RenderOptions, iters: int = 1, trackers: bool = True):
        w, h, intr, c2w = self._cam_args(cam)
        rgba = np.zeros((h, w, 4), np.uint8)
        split = np.zeros((h * w, 3), np.float32) if trackers else None
        sample = np.zeros((h * w, 3), np.float32) if trackers else None
        ms = np.zeros(max(iters, 1), np.float32)
        rc = self.L.ref_render_voxels(self.h, w, h, intr.ctypes.data, c2w.ctypes.data, C.byref(opt),
                                      C.sizeof(opt), rgba.ctypes.data, _ptr(split), _ptr(sample),
                                      iters, ms.ctypes.data)
        assert rc == 0, rc
        return dict(rgba=rgba, to_split=split, to_sample=sample, ms=ms)

    def render_interop(self, cam: dict, opt: RenderOptions, prior_rgba: np.ndarray, depth: np.ndarray,
                       track_visit: bool = False, max_capacity: int = 0):
        """render_voxels(offscreen=false) over a prior colour surface and a mesh-depth surface, like the viewer."""
        w, h, intr, c2w = self._cam_args(cam)
        pr = np.ascontiguousarray(prior_rgba, np.uint8)
        dp = np.ascontiguousarray(depth, np.float32)
        rgba = np.zeros((h, w, 4), np.uint8)
        visited = np.zeros(max_capacity, np.int32) if max_capacity else None
        self.L.ref_render_voxels_interop.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                     C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        rc = self.L.ref_render_voxels_interop(self.h, w, h, intr.ctypes.data, c2w.ctypes.data, C.byref(opt),
                                              C.sizeof(opt), pr.ctypes.data, dp.ctypes.data, rgba.ctypes.data,
                                              int(track_visit), _ptr(visited))
        assert rc == 0, rc
        return rgba, visited

    def render_frame_host(self, cam: dict, opt: RenderOptions, rgba: np.ndarray):
        w, h, intr, c2w = self._cam_args(cam)
        rc = self.L.ref_render_frame_host(self.h, w, h, intr.ctypes.data, c2w.ctypes.data,
                                          C.byref(opt), C.sizeof(opt), rgba.ctypes.data)
        assert rc == 0, rc
        return rgba

    def get_samples(self, cam: dict, opt: RenderOptions, grid_dim, min_position, rng):
        w, h, intr, c2w = self._cam_args(cam)
        P, S, sd = w * h, opt.max_guided_samples, sample_dim(opt)
        ns = np.zeros(P, np.int16)
        samples = np.zeros((P, S, sd), np.float32)
        cluster = np.zeros((P, S), np.int16)
        split = np.zeros((P, 3), np.float32)
        samp = np.zeros((P, 3), np.float32)
        gd = np.ascontiguousarray(grid_dim, np.int32)
        mp = np.ascontiguousarray(min_position, np.float32)
        rg = np.ascontiguousarray(rng, np.float32)
        self.L.ref_get_samples.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_int] + \
            [C.c_void_p] * 3 + [C.c_int, C.c_int] + [C.c_void_p] * 5
        rc = self.L.ref_get_samples(self.h, w, h, intr.ctypes.data, c2w.ctypes.data, C.byref(opt), C.sizeof(opt),
                                    gd.ctypes.data, mp.ctypes.data, rg.ctypes.data, S, sd, ns.ctypes.data,
                                    samples.ctypes.data, cluster.ctypes.data, split.ctypes.data, samp.ctypes.data)
        assert rc == 0, rc
        return dict(num_samples=ns, samples=samples, cluster=cluster, to_split=split, to_sample=samp)

    def render_nerf_results(self, cam: dict, opt: RenderOptions, values, z_vals, offsets):
        w, h, intr, c2w = self._cam_args(cam)
        values = np.ascontiguousarray(values, np.float32)
        z_vals = np.ascontiguousarray(z_vals, np.float32)
        offsets = np.ascontiguousarray(offsets, np.int64)
        rgba = np.zeros((h, w, 4), np.uint8)
        self.L.ref_render_nerf_results.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 3 + \
            [C.c_int, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        rc = self.L.ref_render_nerf_results(self.h, w, h, intr.ctypes.data, c2w.ctypes.data, C.byref(opt),
                                            C.sizeof(opt), values.ctypes.data, values.shape[0], values.shape[1],
                                            z_vals.ctypes.data, offsets.ctypes.data, rgba.ctypes.data)
        if rc == 3:  # the reference's own launch fails on this GPU (168 regs x 512 threads)
            return None
        assert rc == 0, rc
        return rgba

    def add_children(self, opt: RenderOptions, parent_nodes, samples, grid_dim, min_position, rng):
        pn = np.ascontiguousarray(parent_nodes, np.int32)
        sm = np.ascontiguousarray(samples, np.float32).copy()
        n, spc, rd = pn.shape[0], sm.shape[1], sm.shape[2]
        cl = np.zeros((n * 8, spc), np.int16)
        gd, mp, rg = (np.ascontiguousarray(grid_dim, np.int32), np.ascontiguousarray(min_position, np.float32),
                      np.ascontiguousarray(rng, np.float32))
        self.L.ref_add_children.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                            C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        rc = self.L.ref_add_children(self.h, C.byref(opt), C.sizeof(opt), pn.ctypes.data, n, sm.ctypes.data, spc,
                                     rd, cl.ctypes.data, gd.ctypes.data, mp.ctypes.data, rg.ctypes.data)
        assert rc == 0, rc
        return sm, cl

    def generate_samples(self, opt: RenderOptions, nodes, samples, grid_dim, min_position, rng):
        nd = np.ascontiguousarray(nodes, np.int32)
        sm = np.ascontiguousarray(samples, np.float32).copy()
        m, spc, rd = nd.shape[0], sm.shape[1], sm.shape[2]
        cl = np.zeros((m, spc), np.int16)
        gd, mp, rg = (np.ascontiguousarray(grid_dim, np.int32), np.ascontiguousarray(min_position, np.float32),
                      np.ascontiguousarray(rng, np.float32))
        self.L.ref_generate_samples.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                                C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        rc = self.L.ref_generate_samples(self.h, C.byref(opt), C.sizeof(opt), nd.ctypes.data, m, sm.ctypes.data,
                                         spc, rd, cl.ctypes.data, gd.ctypes.data, mp.ctypes.data, rg.ctypes.data)
        assert rc == 0, rc
        return sm, cl

    def prune(self, to_delete):
        td = np.ascontiguousarray(to_delete, np.uint8)
        self.L.ref_prune.argtypes = [C.c_void_p, C.c_void_p]
        rc = self.L.ref_prune(self.h, td.ctypes.data)
        assert rc >= 0, rc
        return rc

    def render_logged(self, cam: dict, opt: RenderOptions, log_cap: int = 0):
        assert self.instr
        w, h, intr, c2w = self._cam_args(cam)
        P = w * h
        rgba = np.zeros((h, w, 4), np.uint8)
        vh = np.zeros(P, np.uint64)
        vc = np.zeros(P, np.int32)
        vlog = np.full((P, log_cap), -1, np.int32) if log_cap > 0 else None
        rc = self.L.ref_render_voxels_logged(self.h, w, h, intr.ctypes.data, c2w.ctypes.data,
                                             C.byref(opt), C.sizeof(opt), rgba.ctypes.data,
                                             vh.ctypes.data, vc.ctypes.data, _ptr(vlog), log_cap)
        assert rc == 0, rc
        return dict(rgba=rgba, hash=vh, count=vc, log=vlog)
