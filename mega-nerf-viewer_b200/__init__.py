"""mega-nerf-viewer_b200 — B200-native octree render path of mega-nerf-viewer.

Python is plumbing only: this module binds ``libmnv_b200.so`` (hand-written
sm_100a CUDA behind the C-ABI declared in ``include/mnv_b200.h``) with ctypes
and uses torch for device buffers / streams.  There is NO fallback: if the
shared library is missing or no CUDA device is present, every compute entry
point raises.

The directory name contains '-', so import it through the root-level shim::

    import mega_nerf_viewer_b200 as mnv
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import synth  # noqa: F401  (synthetic N3Tree generator)
from . import multigpu  # noqa: F401,E402  (sub-module split across GPUs)
from . import export  # noqa: F401,E402  (Mega-NeRF checkpoint -> model container)
from .synth import HostTree

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MNV_B200_LIB") or os.path.join(_HERE, "libmnv_b200.so")  # env override: dev A/B builds

MNV_OK = 0
ERR_NAMES = {1: "INVALID", 2: "NO_DEVICE", 3: "CUDA", 4: "OOM", 5: "IO", 6: "FORMAT", 7: "FULL"}


class MnvError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"mnv_b200 error {code} ({ERR_NAMES.get(code, '?')}): {msg}")
        self.code = code


class RenderOptions(C.Structure):
    """mnv_render_options == viewer::RenderOptions (include/render_options.hpp:9-56)."""

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


class Camera(C.Structure):
    """mnv_camera: CameraSpec (include/data_spec.hpp:9-23) + the 4x3 c2w."""

    _fields_ = [
        ("width", C.c_int), ("height", C.c_int),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("c2w", C.c_float * 12),
    ]


class TreeDesc(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("data_dim", C.c_int), ("format", C.c_int), ("basis_dim", C.c_int),
        ("capacity", C.c_int64),
        ("data", C.c_void_p), ("child", C.c_void_p), ("parent", C.c_void_p),
        ("sample_counts", C.c_void_p),
        ("scale", C.c_float * 3), ("offset", C.c_float * 3),
    ]


class VqDesc(C.Structure):
    _fields_ = [("n_quant", C.c_int), ("n_retain", C.c_int), ("quant_colors", C.c_void_p), ("quant_map", C.c_void_p),
                ("data_retained", C.c_void_p), ("sigma", C.c_void_p)]


class FrameStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("visits", C.c_uint64), ("shaded_visits", C.c_uint64),
                ("rays_hit", C.c_uint64)]


class MlpDesc(C.Structure):
    """mnv_mlp_desc (include/mnv_b200.h)."""

    _fields_ = [
        ("n_trunk_layers", C.c_int), ("width", C.c_int), ("skip_layer", C.c_int),
        ("pe_xyz_freqs", C.c_int), ("pe_dir_freqs", C.c_int), ("need_viewdir", C.c_int),
        ("appearance_dim", C.c_int), ("n_appearance", C.c_int), ("head_width", C.c_int),
        ("out_rgb_dim", C.c_int), ("sigma_activation", C.c_int),
        ("trunk_w", C.c_void_p * 12), ("trunk_b", C.c_void_p * 12),
        ("sigma_w", C.c_void_p), ("sigma_b", C.c_void_p),
        ("final_w", C.c_void_p), ("final_b", C.c_void_p),
        ("embedding", C.c_void_p),
        ("head1_w", C.c_void_p), ("head1_b", C.c_void_p),
        ("head2_w", C.c_void_p), ("head2_b", C.c_void_p),
    ]


_lib = None


def build_library(verbose: bool = False) -> str:
    """Compile csrc/ for sm_100a into libmnv_b200.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libmnv_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stderr[-2000:])
    return LIB_PATH


def lib() -> C.CDLL:
    """The CUDA extension; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU / PyTorch fallback for the render path)")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.mnv_version.restype = C.c_char_p
    L.mnv_last_error.restype = C.c_char_p
    L.mnv_device_count.argtypes = [C.POINTER(C.c_int)]
    L.mnv_render_options_default.argtypes = [C.POINTER(RenderOptions)]
    L.mnv_render_options_default.restype = None
    L.mnv_tree_create.argtypes = [C.POINTER(vp), C.POINTER(TreeDesc), i64, i32]
    L.mnv_tree_create_vq.argtypes = [C.POINTER(vp), C.POINTER(TreeDesc), C.POINTER(VqDesc), i64, i32]
    L.mnv_tree_destroy.argtypes = [vp]
    L.mnv_tree_capacity.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.mnv_tree_device_bytes.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.mnv_tree_download.argtypes = [vp, i64, i64, vp, vp, vp, vp]
    L.mnv_query_points.argtypes = [vp, vp, i64, vp, vp]
    L.mnv_tree_set_tile_order.argtypes = [vp, vp, i32]
    L.mnv_tree_release_surfaces.argtypes = [vp]
    L.mnv_render_voxels.argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderOptions), vp, vp, vp, vp,
                                    vp, vp, C.c_bool, C.c_bool, vp]
    L.mnv_render_voxels_tiles.argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderOptions), vp, vp, vp,
                                          i32, i32, i32, i32, vp]
    L.mnv_render_voxels_logged.argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderOptions), vp, vp, vp,
                                           vp, vp, i32, vp]
    L.mnv_render_frame_host.argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderOptions), vp,
                                        C.POINTER(FrameStats)]
    L.mnv_render_frame_host_bands.argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderOptions), vp,
                                              i32, i32, i32, C.POINTER(FrameStats)]
    L.mnv_tree_trackers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.mnv_guided_samples.argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderOptions), vp, C.c_bool, vp, vp, vp,
                                     vp, vp, vp, i32, vp, i64, C.POINTER(i64), vp, vp, vp, C.c_bool, vp]
    L.mnv_render_nerf_results.argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderOptions), vp, vp, vp, i32, i32,
                                          vp, vp, C.c_bool, vp]
    L.mnv_add_children_and_generate_samples.argtypes = [vp, C.POINTER(RenderOptions), vp, i32, vp, vp, vp, vp, vp, vp, vp]
    L.mnv_tree_commit_children.argtypes = [vp, C.POINTER(RenderOptions), i32, vp, i32, vp]
    L.mnv_tree_record_bytes.argtypes = [vp, C.POINTER(i32)]
    L.mnv_tree_reduce_children.argtypes = [vp, C.POINTER(RenderOptions), i32, vp, i32, vp, vp]
    L.mnv_tree_commit_children_records.argtypes = [vp, C.POINTER(RenderOptions), i32, vp, vp]
    L.mnv_generate_samples.argtypes = [vp, C.POINTER(RenderOptions), vp, i32, vp, vp, vp, vp, vp, vp]
    L.mnv_tree_update_samples.argtypes = [vp, C.POINTER(RenderOptions), vp, i32, vp, i32, vp]
    L.mnv_tree_prune.argtypes = [vp, vp, vp, i32, i64, vp]
    L.mnv_query_submodules.argtypes = [vp, vp, vp, i32, i64, vp, i32, vp]
    L.mnv_select_split_candidates.argtypes = [vp, i64, i32, vp, C.POINTER(i32), C.POINTER(i32), vp]
    L.mnv_select_sample_candidates.argtypes = [vp, i64, i32, vp, C.POINTER(i32), C.POINTER(i32), vp]
    L.mnv_vote_reduce.argtypes = [vp, i64, vp, i64, C.POINTER(i64), vp]
    L.mnv_select_candidates_from_votes.argtypes = [i32, vp, i64, vp, i64, i32, vp, C.POINTER(i32), C.POINTER(i32), vp]
    L.mnv_tracker_encode_chunk.argtypes = [i32]
    L.mnv_tracker_encode_chunk.restype = C.c_float
    L.mnv_tracker_decode_chunk.argtypes = [C.c_float]
    L.mnv_tracker_decode_chunk.restype = i32
    L.mnv_group_create.argtypes = [C.POINTER(vp), C.POINTER(TreeDesc), i64, C.POINTER(i32), i32]
    L.mnv_group_destroy.argtypes = [vp]
    L.mnv_group_size.argtypes = [vp, C.POINTER(i32)]
    L.mnv_group_tree.argtypes = [vp, i32, C.POINTER(vp)]
    L.mnv_group_render_frame.argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderOptions), vp, vp, i32]
    L.mnv_group_render_frame_host.argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderOptions), vp, i32]
    L.mnv_group_synchronize.argtypes = [vp]
    L.mnv_group_refine_frame.argtypes = [vp, C.POINTER(vp), C.POINTER(Camera), C.POINTER(RenderOptions), vp, vp, vp,
                                         C.c_uint64, vp, vp, vp, i32, C.POINTER(i32)]
    L.mnv_model_create.argtypes = [C.POINTER(vp), i32, C.POINTER(MlpDesc), vp, vp, vp, i32]
    L.mnv_model_destroy.argtypes = [vp]
    L.mnv_model_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_double)]
    L.mnv_mlp_forward.argtypes = [vp, i32, vp, i64, i32, vp, i32, vp]
    L.mnv_render_voxels_partial.argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderOptions), vp, i32, C.POINTER(vp),
                                            i32, i32, vp]
    L.mnv_signal_peers.argtypes = [C.POINTER(vp), i32, i32, C.c_uint32, vp]
    L.mnv_composite_partials.argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderOptions), vp, i32, i32, vp, i64, i32,
                                         vp, vp, C.c_uint32, vp]
    L.mnv_composite_partials_guided.argtypes = L.mnv_composite_partials.argtypes
    L.mnv_guided_segment_probe.argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderOptions), vp, vp, vp]
    L.mnv_guided_samples_segment.argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderOptions), vp, vp, vp, vp, vp, i32, i32,
                                             vp, vp, vp, i32, vp, i64, C.POINTER(i64), vp]
    L.mnv_render_nerf_results_partial.argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderOptions), vp, i32, i32, vp, vp,
                                                  vp, i32, i32, i32, C.POINTER(vp), i32, vp]
    L.mnv_ipc_export.argtypes = [vp, C.c_char_p]
    L.mnv_ipc_open.argtypes = [C.c_char_p, C.POINTER(vp), i32]
    L.mnv_ipc_close.argtypes = [vp]
    L.mnv_array_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32]
    L.mnv_array_destroy.argtypes = [vp]
    L.mnv_array_upload.argtypes = [vp, vp, C.c_size_t, i32]
    L.mnv_array_download.argtypes = [vp, vp, C.c_size_t, i32]
    L.mnv_malloc.argtypes = [C.POINTER(vp), C.c_size_t, i32]
    L.mnv_free.argtypes = [vp]
    L.mnv_memset.argtypes = [vp, i32, C.c_size_t, vp]
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != MNV_OK:
        raise MnvError(rc, lib().mnv_last_error().decode(errors="replace"))


def device_count() -> int:
    n = C.c_int(0)
    _check(lib().mnv_device_count(C.byref(n)))
    return n.value


def default_options(**kw) -> RenderOptions:
    """RenderOptions with the struct defaults of include/render_options.hpp."""
    o = RenderOptions()
    lib().mnv_render_options_default(C.byref(o))
    for k, v in kw.items():
        if isinstance(v, (list, tuple, np.ndarray)):
            getattr(o, k)[:] = list(v)
        else:
            setattr(o, k, v)
    return o


def make_camera(d) -> Camera:
    if isinstance(d, Camera):
        return d
    c = Camera()
    c.width, c.height = int(d["width"]), int(d["height"])
    c.fx, c.fy, c.cx, c.cy = d["fx"], d["fy"], d["cx"], d["cy"]
    c.c2w[:] = [float(v) for v in d["c2w"]]
    return c


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise MnvError(2, "torch sees no CUDA device; the render path has no CPU fallback")
    return torch


def _dptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(stream):
    if stream is None:
        import torch

        stream = torch.cuda.current_stream()
    return C.c_void_p(stream.cuda_stream)


class DeviceTree:
    """Device-resident N3Tree in the SoA layout (csrc/mnv_internal.cuh)."""

    def __init__(self, tree: HostTree, max_capacity: int = 0, device: int = 0,
                 sample_counts: np.ndarray | None = None, vq: dict | None = None):
        """vq: dict(quant_colors f16 [n_q, 65536, 3], quant_map u16 [n_q, cap, 8], data_retained f16 [n_r, cap, 8, 3] or
        None, sigma f16 [cap, 8]) -> the leaf payloads are decoded on the device (mnv_tree_create_vq); tree.data is
        not read."""
        d = TreeDesc()
        d.N = tree.N
        d.data_dim = tree.data_dim
        d.format = 1 if tree.data_format.upper().startswith("SH") else 0
        d.basis_dim = tree.basis_dim
        d.capacity = tree.capacity
        self._keep = [np.ascontiguousarray(tree.data.view(np.uint16)),
                      np.ascontiguousarray(tree.child, np.int32),
                      np.ascontiguousarray(tree.parent, np.int32)]
        d.data, d.child, d.parent = (a.ctypes.data for a in self._keep)
        if sample_counts is not None:
            sc = np.ascontiguousarray(sample_counts, np.int16)
            self._keep.append(sc)
            d.sample_counts = sc.ctypes.data
        d.scale[:] = [float(v) for v in tree.scale]
        d.offset[:] = [float(v) for v in tree.offset]
        self.data_dim = tree.data_dim
        self.device = device
        self._h = C.c_void_p()
        if vq is not None:
            q = VqDesc()
            book = np.ascontiguousarray(vq["quant_colors"]).view(np.uint16)
            qmap = np.ascontiguousarray(vq["quant_map"], np.uint16)
            ret = vq.get("data_retained")
            ret = None if ret is None or len(ret) == 0 else np.ascontiguousarray(ret).view(np.uint16)
            sig = np.ascontiguousarray(vq["sigma"]).view(np.uint16)
            self._keep += [book, qmap, ret, sig]
            q.n_quant, q.n_retain = qmap.shape[0], 0 if ret is None else ret.shape[0]
            q.quant_colors, q.quant_map, q.sigma = book.ctypes.data, qmap.ctypes.data, sig.ctypes.data
            q.data_retained = None if ret is None else ret.ctypes.data
            d.data = None
            _check(lib().mnv_tree_create_vq(C.byref(self._h), C.byref(d), C.byref(q), max_capacity, device))
        else:
            _check(lib().mnv_tree_create(C.byref(self._h), C.byref(d), max_capacity, device))
        self._keep = None  # host arrays are not referenced after the upload

    def close(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value and _lib is not None and getattr(self, "_owned", True):
            _lib.mnv_tree_destroy(h)
        self._h = None

    __del__ = close

    @property
    def capacity(self) -> int:
        a, b = C.c_int64(), C.c_int64()
        _check(lib().mnv_tree_capacity(self._h, C.byref(a), C.byref(b)))
        return a.value

    @property
    def max_capacity(self) -> int:
        a, b = C.c_int64(), C.c_int64()
        _check(lib().mnv_tree_capacity(self._h, C.byref(a), C.byref(b)))
        return b.value

    def device_bytes(self) -> int:
        n = C.c_uint64()
        _check(lib().mnv_tree_device_bytes(self._h, C.byref(n)))
        return n.value

    def download(self, first: int = 0, count: int | None = None):
        """Back to the reference's AoS arrays: (data f16, child, parent, sample_counts)."""
        count = self.capacity - first if count is None else count
        data = np.empty((count, 8, self.data_dim), np.uint16)
        child = np.empty((count, 8), np.int32)
        parent = np.empty(count, np.int32)
        sc = np.empty((count, 8), np.int16)
        _check(lib().mnv_tree_download(self._h, first, count, data.ctypes.data, child.ctypes.data,
                                       parent.ctypes.data, sc.ctypes.data))
        return data.view(np.float16), child, parent, sc

    def set_tile_order(self, order=None):
        """order: CUDA int32 tensor, a permutation of the frame's 16x8 CTA tiles (kept alive by this object)."""
        self._tile_order = order
        _check(lib().mnv_tree_set_tile_order(self._h, _dptr(order), 0 if order is None else order.numel()))

    # ---- point query (rt_core.cuh:117-159) ---------------------------------
    def query_points(self, xyz, stream=None):
        torch = _torch()
        xyz = torch.as_tensor(xyz, dtype=torch.float32, device=f"cuda:{self.device}").contiguous()
        out = torch.empty((xyz.shape[0], 3), dtype=torch.int32, device=xyz.device)
        _check(lib().mnv_query_points(self._h, _dptr(xyz), xyz.shape[0], _dptr(out), _stream_ptr(stream)))
        return out

    # ---- octree render (renderer_kernel.cu:396-437) --------------------------
    def render(self, cam, opt: RenderOptions, *, out=None, to_split=None, to_sample=None,
               visited=None, track_visit=False, stream=None):
        """Offscreen render into a device RGBA8 tensor [H, W, 4]."""
        torch = _torch()
        cam = make_camera(cam)
        if out is None:
            out = torch.empty((cam.height, cam.width, 4), dtype=torch.uint8, device=f"cuda:{self.device}")
        _check(lib().mnv_render_voxels(self._h, C.byref(cam), C.byref(opt), None, None, _dptr(out),
                                       _dptr(to_split), _dptr(to_sample), _dptr(visited),
                                       track_visit, True, _stream_ptr(stream)))
        return out

    def render_tiles(self, cam, opt, out, tile_w, tile_h, tile_mod, tile_rem, *, to_split=None,
                     to_sample=None, stream=None):
        cam = make_camera(cam)
        _check(lib().mnv_render_voxels_tiles(self._h, C.byref(cam), C.byref(opt), _dptr(out),
                                             _dptr(to_split), _dptr(to_sample), tile_w, tile_h,
                                             tile_mod, tile_rem, _stream_ptr(stream)))
        return out

    def render_logged(self, cam, opt, log_cap: int = 0, stream=None):
        """Parity helper: image + per-ray visit hash / count / shaded count / log."""
        torch = _torch()
        cam = make_camera(cam)
        dev = f"cuda:{self.device}"
        P = cam.width * cam.height
        img = torch.empty((cam.height, cam.width, 4), dtype=torch.uint8, device=dev)
        vh = torch.zeros(P, dtype=torch.int64, device=dev)
        vc = torch.zeros(P, dtype=torch.int32, device=dev)
        vs = torch.zeros(P, dtype=torch.int32, device=dev)
        vlog = torch.full((P, log_cap), -1, dtype=torch.int32, device=dev) if log_cap > 0 else None
        _check(lib().mnv_render_voxels_logged(self._h, C.byref(cam), C.byref(opt), _dptr(img), _dptr(vh),
                                              _dptr(vc), _dptr(vs), _dptr(vlog), log_cap,
                                              _stream_ptr(stream)))
        torch.cuda.synchronize()
        return dict(rgba=img.cpu().numpy(), hash=vh.cpu().numpy().view(np.uint64),
                    count=vc.cpu().numpy(), shaded=vs.cpu().numpy(),
                    log=None if vlog is None else vlog.cpu().numpy())

    # ---- guided sampling (renderer_kernel.cu:439-485 + cuda_renderer.cpp:116-120) ----------
    def guided_samples(self, cam, opt: RenderOptions, grid_dim, min_position, rng, capacity_rows: int,
                       to_split=None, to_sample=None, stream=None):
        """-> dict(total, offsets i64 [P], z_vals [V], rows [V, in_dim], cluster i16 [V]) on the device."""
        torch = _torch()
        cam = make_camera(cam)
        dev = f"cuda:{self.device}"
        P = cam.width * cam.height
        in_dim = 3 + (3 if opt.need_viewdir else 0) + (1 if opt.appearance_embedding != -1 else 0)
        offsets = torch.empty(P, dtype=torch.int64, device=dev)
        z = torch.empty(capacity_rows, dtype=torch.float32, device=dev)
        rows = torch.empty((capacity_rows, in_dim), dtype=torch.float32, device=dev)
        cluster = torch.empty(capacity_rows, dtype=torch.int16, device=dev)
        gd = np.ascontiguousarray(grid_dim, np.int32)
        mp = np.ascontiguousarray(min_position, np.float32)
        rg = np.ascontiguousarray(rng, np.float32)
        total = C.c_int64(0)
        _check(lib().mnv_guided_samples(self._h, C.byref(cam), C.byref(opt), None, True, gd.ctypes.data,
                                        mp.ctypes.data, rg.ctypes.data, _dptr(offsets), _dptr(z), _dptr(rows),
                                        in_dim, _dptr(cluster), capacity_rows, C.byref(total), _dptr(to_split),
                                        _dptr(to_sample), None, False, _stream_ptr(stream)))
        v = total.value
        return dict(total=v, offsets=offsets, z_vals=z[:v], rows=rows[:v], cluster=cluster[:v])

    def render_nerf_results(self, cam, opt: RenderOptions, values, z_vals, offsets, sigma_col: int = -1,
                            out=None, stream=None):
        """Composite MLP outputs (values f32 [V, stride]) per ray -> RGBA8 [H, W, 4] on the device."""
        torch = _torch()
        cam = make_camera(cam)
        if out is None:
            out = torch.empty((cam.height, cam.width, 4), dtype=torch.uint8, device=f"cuda:{self.device}")
        _check(lib().mnv_render_nerf_results(self._h, C.byref(cam), C.byref(opt), None, _dptr(out), _dptr(values),
                                             values.stride(0), sigma_col, _dptr(z_vals), _dptr(offsets), True,
                                             _stream_ptr(stream)))
        return out

    # ---- refinement (renderer_kernel.cu:63-213, cuda_renderer.cpp:205-381) --------------------
    @staticmethod
    def _grid_args(grid_dim, min_position, rng):
        a = (np.ascontiguousarray(grid_dim, np.int32), np.ascontiguousarray(min_position, np.float32),
             np.ascontiguousarray(rng, np.float32))
        return a, [C.c_void_p(x.ctypes.data) for x in a]

    def add_children(self, opt, parent_nodes, samples, cluster, grid_dim, min_position, rng, visited=None,
                     stream=None):
        keep, g = self._grid_args(grid_dim, min_position, rng)
        _check(lib().mnv_add_children_and_generate_samples(
            self._h, C.byref(opt), _dptr(parent_nodes), parent_nodes.shape[0], _dptr(samples), _dptr(cluster),
            _dptr(visited), g[0], g[1], g[2], _stream_ptr(stream)))

    def commit_children(self, opt, n, results, stream=None):
        _check(lib().mnv_tree_commit_children(self._h, C.byref(opt), n, _dptr(results), results.shape[-1],
                                              _stream_ptr(stream)))

    @property
    def record_bytes(self) -> int:
        b = C.c_int(0)
        _check(lib().mnv_tree_record_bytes(self._h, C.byref(b)))
        return b.value

    def reduce_children(self, opt, results, stream=None):
        """results f32 [n_children, samples_per_corner, stride] -> payload records uint8 [n_children, record_bytes]."""
        torch = _torch()
        n = results.shape[0]
        rec = torch.empty((n, self.record_bytes), dtype=torch.uint8, device=results.device)
        _check(lib().mnv_tree_reduce_children(self._h, C.byref(opt), n, _dptr(results), results.shape[-1], _dptr(rec),
                                              _stream_ptr(stream)))
        return rec

    def commit_children_records(self, opt, n, records, stream=None):
        assert records.shape[0] >= n * 8 and records.is_contiguous()
        _check(lib().mnv_tree_commit_children_records(self._h, C.byref(opt), n, _dptr(records), _stream_ptr(stream)))

    def generate_samples(self, opt, nodes, samples, cluster, grid_dim, min_position, rng, stream=None):
        keep, g = self._grid_args(grid_dim, min_position, rng)
        _check(lib().mnv_generate_samples(self._h, C.byref(opt), _dptr(nodes), nodes.shape[0], _dptr(samples),
                                          _dptr(cluster), g[0], g[1], g[2], _stream_ptr(stream)))

    def update_samples(self, opt, nodes, results, stream=None):
        _check(lib().mnv_tree_update_samples(self._h, C.byref(opt), _dptr(nodes), nodes.shape[0],
                                             _dptr(results), results.shape[-1], _stream_ptr(stream)))

    def prune(self, to_delete, stream=None):
        """to_delete: CUDA bool/uint8 [capacity]. Mirrors Impl::prune_tree (cumsum on the device)."""
        torch = _torch()
        td = to_delete.to(torch.uint8).contiguous()
        shifts = torch.cumsum(td, 0, dtype=torch.int32)
        num = int(shifts[-1].item())
        if num == 0:
            return 0
        first = int(torch.argmin(shifts).item())  # cuda_renderer.cpp:357
        _check(lib().mnv_tree_prune(self._h, _dptr(td), _dptr(shifts), first, num, _stream_ptr(stream)))
        return num

    # ---- the viewer's presentation path: cudaArray surfaces, offscreen = false -----------------
    def render_interop(self, cam, opt: RenderOptions, prior_rgba: np.ndarray, depth: np.ndarray,
                       track_visit: bool = False, visited=None, to_split=None, to_sample=None):
        """mnv_render_voxels exactly as Impl::render drives it (cuda_renderer.cpp:141-142): the RGBA8 surface
        already holds what GL drew (`prior_rgba` [H, W, 4] u8), the R32F surface the mesh depth (`depth`
        [H, W] f32), the frame is composited over them in place.  Returns the RGBA8 frame (numpy)."""
        torch = _torch()
        cam = make_camera(cam)
        h, w = cam.height, cam.width
        img, dep = C.c_void_p(), C.c_void_p()
        _check(lib().mnv_array_create(C.byref(img), w, h, 0, self.device))
        _check(lib().mnv_array_create(C.byref(dep), w, h, 1, self.device))
        try:
            pr = np.ascontiguousarray(prior_rgba, np.uint8)
            dp = np.ascontiguousarray(depth, np.float32)
            _check(lib().mnv_array_upload(img, pr.ctypes.data, w * 4, h))
            _check(lib().mnv_array_upload(dep, dp.ctypes.data, w * 4, h))
            _check(lib().mnv_render_voxels(self._h, C.byref(cam), C.byref(opt), img, dep, None, _dptr(to_split),
                                           _dptr(to_sample), _dptr(visited), track_visit, False, _stream_ptr(None)))
            torch.cuda.synchronize()
            out = np.empty((h, w, 4), np.uint8)
            _check(lib().mnv_array_download(out.ctypes.data, img, w * 4, h))
        finally:
            lib().mnv_tree_release_surfaces(self._h)  # the arrays die here: drop the surface objects bound to them
            lib().mnv_array_destroy(img)
            lib().mnv_array_destroy(dep)
        return out

    # ---- sub-module split across GPUs (csrc/mnv_multigpu.cu) ---------------------------------
    def render_partial(self, cam, opt: RenderOptions, dst_ptrs, block_pixels: int, slot: int, cell_box=None,
                       stream=None):
        """March the frame clipped to opt.render_bbox and `cell_box` (tree space) and store each ray's premultiplied
        (r, g, b, alpha) into the owners' buffers (`dst_ptrs`: device addresses, local or IPC-mapped peers)."""
        cam = make_camera(cam)
        arr = (C.c_void_p * len(dst_ptrs))(*[C.c_void_p(int(a)) for a in dst_ptrs])
        cb = None if cell_box is None else np.ascontiguousarray(cell_box, np.float32)
        _check(lib().mnv_render_voxels_partial(self._h, C.byref(cam), C.byref(opt), None if cb is None else cb.ctypes.data,
                                               len(dst_ptrs), arr, block_pixels, slot, _stream_ptr(stream)))

    def guided_segment_probe(self, cam, opt: RenderOptions, cell_box, out=None, stream=None):
        """Probe pass of the sharded guided frame -> f32 [P, 4] = (T at the cell's exit, samples, first z, 0)."""
        torch = _torch()
        cam = make_camera(cam)
        if out is None:
            out = torch.empty((cam.width * cam.height, 4), dtype=torch.float32, device=f"cuda:{self.device}")
        cb = np.ascontiguousarray(cell_box, np.float32)
        _check(lib().mnv_guided_segment_probe(self._h, C.byref(cam), C.byref(opt), cb.ctypes.data, _dptr(out),
                                              _stream_ptr(stream)))
        return out

    def guided_samples_segment(self, cam, opt: RenderOptions, cell_box, grid_dim, min_position, rng, probe_all,
                               slot: int, capacity_rows: int, stream=None):
        """guided_samples for this rank's segment; probe_all f32 [n_cells, P, 4] (all ranks' probe records)."""
        torch = _torch()
        cam = make_camera(cam)
        dev = f"cuda:{self.device}"
        P = cam.width * cam.height
        in_dim = 3 + (3 if opt.need_viewdir else 0) + (1 if opt.appearance_embedding != -1 else 0)
        offsets = torch.empty(P, dtype=torch.int64, device=dev)
        z = torch.empty(capacity_rows, dtype=torch.float32, device=dev)
        rows = torch.empty((capacity_rows, in_dim), dtype=torch.float32, device=dev)
        cluster = torch.empty(capacity_rows, dtype=torch.int16, device=dev)
        gd = np.ascontiguousarray(grid_dim, np.int32)
        mp = np.ascontiguousarray(min_position, np.float32)
        rg = np.ascontiguousarray(rng, np.float32)
        total = C.c_int64(0)
        assert probe_all.is_contiguous() and probe_all.shape[1:] == (P, 4)
        cb = np.ascontiguousarray(cell_box, np.float32)
        _check(lib().mnv_guided_samples_segment(self._h, C.byref(cam), C.byref(opt), cb.ctypes.data, gd.ctypes.data,
                                                mp.ctypes.data,
                                                rg.ctypes.data, _dptr(probe_all), probe_all.shape[0], slot,
                                                _dptr(offsets), _dptr(z), _dptr(rows), in_dim, _dptr(cluster),
                                                capacity_rows, C.byref(total), _stream_ptr(stream)))
        v = total.value
        return dict(total=v, offsets=offsets, z_vals=z[:v], rows=rows[:v], cluster=cluster[:v])

    def render_nerf_results_partial(self, cam, opt: RenderOptions, values, z_vals, offsets, probe_all, slot: int,
                                    dst_ptrs, block_pixels: int, sigma_col: int = -1, stream=None):
        cam = make_camera(cam)
        arr = (C.c_void_p * len(dst_ptrs))(*[C.c_void_p(int(a)) for a in dst_ptrs])
        _check(lib().mnv_render_nerf_results_partial(self._h, C.byref(cam), C.byref(opt), _dptr(values),
                                                     values.stride(0), sigma_col, _dptr(z_vals), _dptr(offsets),
                                                     _dptr(probe_all), probe_all.shape[0], slot, len(dst_ptrs), arr,
                                                     block_pixels, _stream_ptr(stream)))

    def composite_partials(self, cam, opt: RenderOptions, partials_ptr: int, n: int, block_pixels: int, boxes,
                           first_pixel: int, n_pixels: int, out, flags_ptr: int = 0, wait_value: int = 0, stream=None,
                           guided: bool = False):
        cam = make_camera(cam)
        bx = np.ascontiguousarray(boxes, np.float32)
        fn = lib().mnv_composite_partials_guided if guided else lib().mnv_composite_partials
        _check(fn(self._h, C.byref(cam), C.byref(opt), C.c_void_p(partials_ptr), n,
                                            block_pixels, bx.ctypes.data, first_pixel, n_pixels, _dptr(out),
                                            C.c_void_p(flags_ptr) if flags_ptr else None, wait_value,
                                            _stream_ptr(stream)))
        return out

    def render_frame_host(self, cam, opt, rgba_host=None, stats: bool = False, bands=None):
        """The per-frame call with HOST buffers (camera in, RGBA8 frame out).
        rgba_host: numpy array or a (pinned) torch CPU tensor [H, W, 4] uint8.
        bands=(band_rows, mod, rem): multi-GPU band partition of the frame."""
        cam = make_camera(cam)
        if rgba_host is None:
            rgba_host = np.empty((cam.height, cam.width, 4), np.uint8)
        ptr = rgba_host.ctypes.data if isinstance(rgba_host, np.ndarray) else rgba_host.data_ptr()
        st = FrameStats() if stats else None
        if bands is None:
            _check(lib().mnv_render_frame_host(self._h, C.byref(cam), C.byref(opt), C.c_void_p(ptr),
                                               C.byref(st) if stats else None))
        else:
            _check(lib().mnv_render_frame_host_bands(self._h, C.byref(cam), C.byref(opt),
                                                     C.c_void_p(ptr), bands[0], bands[1], bands[2],
                                                     C.byref(st) if stats else None))
        if stats:
            return rgba_host, dict(rays=st.rays, visits=st.visits, shaded_visits=st.shaded_visits,
                                   rays_hit=st.rays_hit)
        return rgba_host


class ReplicaGroup:
    """mnv_group: the tree replicated on several GPUs of one box, driven from this one process (image tiles:
    interleaved bands per replica, NVLink peer copies into devices[0]; refinement with vote / payload records
    exchanged by peer copies).  devices may repeat (tests on one GPU)."""

    def __init__(self, tree: HostTree, devices, max_capacity: int = 0):
        d = TreeDesc()
        d.N, d.data_dim = tree.N, tree.data_dim
        d.format = 1 if tree.data_format.upper().startswith("SH") else 0
        d.basis_dim, d.capacity = tree.basis_dim, tree.capacity
        keep = [np.ascontiguousarray(tree.data.view(np.uint16)), np.ascontiguousarray(tree.child, np.int32),
                np.ascontiguousarray(tree.parent, np.int32)]
        d.data, d.child, d.parent = (a.ctypes.data for a in keep)
        d.scale[:] = [float(v) for v in tree.scale]
        d.offset[:] = [float(v) for v in tree.offset]
        self.devices = [int(x) for x in devices]
        self.data_dim = tree.data_dim
        dv = (C.c_int32 * len(self.devices))(*self.devices)
        self._h = C.c_void_p()
        _check(lib().mnv_group_create(C.byref(self._h), C.byref(d), max_capacity, dv, len(self.devices)))

    def close(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value and _lib is not None:
            _lib.mnv_group_destroy(h)
        self._h = None

    __del__ = close

    def replica(self, i: int) -> "DeviceTree":
        """Non-owning DeviceTree view of replica i (download / capacity)."""
        t = DeviceTree.__new__(DeviceTree)
        h = C.c_void_p()
        _check(lib().mnv_group_tree(self._h, i, C.byref(h)))
        t._h, t.device, t.data_dim, t._keep = h, self.devices[i], self.data_dim, None
        t._owned = False  # the group destroys its replicas
        return t

    def render_frame_host(self, cam, opt, band_rows: int = 8) -> np.ndarray:
        cam = make_camera(cam)
        out = np.empty((cam.height, cam.width, 4), np.uint8)
        _check(lib().mnv_group_render_frame_host(self._h, C.byref(cam), C.byref(opt), out.ctypes.data, band_rows))
        return out

    def render_frame(self, cam, opt, out, band_rows: int = 8):
        """out: CUDA uint8 [H, W, 4] on devices[0]; asynchronous (ordered on replica 0's stream)."""
        cam = make_camera(cam)
        _check(lib().mnv_group_render_frame(self._h, C.byref(cam), C.byref(opt), _dptr(out), None, band_rows))

    def synchronize(self):
        _check(lib().mnv_group_synchronize(self._h))

    def refine_frame(self, models, cam, opt, grid_dim, min_position, rng, seed: int = 0x5eed, band_rows: int = 8,
                     want_frame: bool = True):
        """models: one MlpModel per replica (on that replica's device) -> (RGBA8 frame or None, leaves split)."""
        cam = make_camera(cam)
        out = np.empty((cam.height, cam.width, 4), np.uint8) if want_frame else None
        keep, g = DeviceTree._grid_args(grid_dim, min_position, rng)
        mh = (C.c_void_p * len(models))(*[m._h for m in models])
        k = C.c_int(0)
        _check(lib().mnv_group_refine_frame(self._h, mh, C.byref(cam), C.byref(opt), g[0], g[1], g[2], seed, None, None,
                                            None if out is None else out.ctypes.data, band_rows, C.byref(k)))
        return out, k.value


def select_candidates(tracker, max_n: int, kind: str = "split", stream=None):
    """tracker: CUDA f32 [P, 3] = (priority, chunk, child) -> (nodes i32 [n, 2] on device, n_candidates)."""
    torch = _torch()
    nodes = torch.empty((max_n, 2), dtype=torch.int32, device=tracker.device)
    n, nc = C.c_int(0), C.c_int(0)
    fn = lib().mnv_select_split_candidates if kind == "split" else lib().mnv_select_sample_candidates
    _check(fn(_dptr(tracker), tracker.shape[0], max_n, _dptr(nodes), C.byref(n), C.byref(nc), _stream_ptr(stream)))
    return nodes[: n.value], nc.value


def vote_reduce(tracker, cap_records: int | None = None, stream=None):
    """tracker: CUDA f32 [P, 3] -> vote records u32 [n, 3] = (leaf id, priority, votes) on the device (unordered)."""
    torch = _torch()
    cap = int(cap_records or tracker.shape[0])
    rec = torch.empty((cap, 3), dtype=torch.int32, device=tracker.device)
    n = C.c_int64(0)
    _check(lib().mnv_vote_reduce(_dptr(tracker), tracker.shape[0], _dptr(rec), cap, C.byref(n), _stream_ptr(stream)))
    return rec[: n.value]


def select_from_votes(records, max_n: int, kind: str = "split", tracker=None, stream=None):
    """Selection over gathered vote records (i32/u32 [n, 3], votes == 0 rows are padding) and optional local rows."""
    torch = _torch()
    nodes = torch.empty((max_n, 2), dtype=torch.int32, device=records.device)
    n, nc = C.c_int(0), C.c_int(0)
    _check(lib().mnv_select_candidates_from_votes(
        0 if kind == "split" else 1, _dptr(tracker) if tracker is not None else None,
        tracker.shape[0] if tracker is not None else 0, _dptr(records), records.shape[0], max_n, _dptr(nodes),
        C.byref(n), C.byref(nc), _stream_ptr(stream)))
    return nodes[: n.value], nc.value


def decode_tracker_chunks(col: np.ndarray) -> np.ndarray:
    """Chunk column of a tracker (f32) -> node ids (i64; -1 = none), see mnv_tracker_decode_chunk."""
    col = np.ascontiguousarray(col, np.float32)
    bits = col.view(np.int32).astype(np.int64)
    big = (bits >= (1 << 24)) & (bits < (1 << 28))
    return np.where(big, bits, np.where(col >= 0, col, -1).astype(np.int64))


class MlpModel:
    """Mega-NeRF sub-MLP container on the device (mnv_model_create / mnv_mlp_forward).

    `submodules`: list of dicts with numpy fp32 arrays in nn.Linear layout:
      trunk_w[i] [256, in], trunk_b[i], sigma_w [1,256], sigma_b [1], final_w/b,
      embedding [n_app, app_dim] (optional), head1_w/b, head2_w/b, and the ints
      skip_layer, pe_xyz_freqs, pe_dir_freqs, need_viewdir, sigma_activation.
    """

    def __init__(self, submodules, grid_dim=(1, 1), min_position=(0, 0, 0), max_position=(1, 1, 1),
                 device: int = 0):
        descs = (MlpDesc * len(submodules))()
        keep = []

        def ptr(a):
            a = np.ascontiguousarray(a, np.float32)
            keep.append(a)
            return a.ctypes.data

        for d, sm in zip(descs, submodules):
            n = len(sm["trunk_w"])
            d.n_trunk_layers, d.width = n, int(sm["trunk_w"][0].shape[0])
            d.skip_layer = int(sm.get("skip_layer", 4))
            d.pe_xyz_freqs = int(sm.get("pe_xyz_freqs", 12))
            d.pe_dir_freqs = int(sm.get("pe_dir_freqs", 4))
            d.need_viewdir = int(bool(sm.get("need_viewdir", False)))
            emb = sm.get("embedding")
            d.appearance_dim = 0 if emb is None else int(emb.shape[1])
            d.n_appearance = 0 if emb is None else int(emb.shape[0])
            d.head_width = int(sm["head1_w"].shape[0])
            d.out_rgb_dim = int(sm["head2_w"].shape[0])
            d.sigma_activation = int(sm.get("sigma_activation", 1))
            for i in range(n):
                d.trunk_w[i], d.trunk_b[i] = ptr(sm["trunk_w"][i]), ptr(sm["trunk_b"][i])
            d.sigma_w, d.sigma_b = ptr(sm["sigma_w"]), ptr(sm["sigma_b"])
            d.final_w, d.final_b = ptr(sm["final_w"]), ptr(sm["final_b"])
            d.embedding = None if emb is None else ptr(emb)
            d.head1_w, d.head1_b = ptr(sm["head1_w"]), ptr(sm["head1_b"])
            d.head2_w, d.head2_b = ptr(sm["head2_w"]), ptr(sm["head2_b"])
        gd = np.asarray(grid_dim, np.int32)
        mn = np.asarray(min_position, np.float32)
        mx = np.asarray(max_position, np.float32)
        self._h = C.c_void_p()
        self.device = device
        _check(lib().mnv_model_create(C.byref(self._h), len(submodules), descs, gd.ctypes.data,
                                      mn.ctypes.data, mx.ctypes.data, device))
        n, i, o, f = C.c_int(), C.c_int(), C.c_int(), C.c_double()
        _check(lib().mnv_model_info(self._h, C.byref(n), C.byref(i), C.byref(o), C.byref(f)))
        self.n_submodules, self.in_dim, self.out_dim, self.flops_per_row = n.value, i.value, o.value, f.value

    def close(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value and _lib is not None:
            _lib.mnv_model_destroy(h)
        self._h = None

    __del__ = close

    def query_submodules(self, cluster, rows, out, stream=None):
        """cluster i16 [V], rows f32 [V, in_dim], out f32 [V, >= out_dim] (all CUDA)."""
        _check(lib().mnv_query_submodules(self._h, _dptr(cluster), _dptr(rows), rows.shape[1], rows.shape[0],
                                          _dptr(out), out.stride(0), _stream_ptr(stream)))
        return out

    def forward(self, x, submodule: int = 0, out=None, stream=None):
        """x: CUDA float32 [rows, in_dim] -> CUDA float32 [rows, out_dim]."""
        torch = _torch()
        x = x.contiguous()
        assert x.dtype == torch.float32 and x.is_cuda and x.shape[1] == self.in_dim
        if out is None:
            out = torch.empty((x.shape[0], self.out_dim), dtype=torch.float32, device=x.device)
        _check(lib().mnv_mlp_forward(self._h, submodule, _dptr(x), x.shape[0], x.shape[1], _dptr(out),
                                     out.stride(0), _stream_ptr(stream)))
        return out


def save_model_container(path: str, submodules, grid_dim=(1, 1), min_position=(0, 0, 0),
                         max_position=(1, 1, 1), centroids=None, need_viewdir: bool | None = None,
                         need_appearance_embedding: bool | None = None, compressed: bool = False) -> None:
    """Write the Mega-NeRF sub-module container the C++ viewer::VolumeRenderer::load_model reads
    (csrc/viewer/model.hpp): the attributes the reference pulls out of its TorchScript archive
    (cuda_renderer.cpp:518-543) plus the flat fp32 weights of every sub_module_<i>.
    `submodules`: the dicts MlpModel takes (tests/mlp_reference.MegaNerfMLP.export())."""
    n = len(submodules)
    if centroids is None:
        centroids = np.zeros((n, 3), np.float32)
    if need_viewdir is None:
        need_viewdir = bool(submodules[0].get("need_viewdir", False))
    if need_appearance_embedding is None:
        need_appearance_embedding = submodules[0].get("embedding") is not None
    arrays = dict(grid_dim=np.asarray(grid_dim, np.int32), min_position=np.asarray(min_position, np.float32),
                  max_position=np.asarray(max_position, np.float32),
                  centroids=np.asarray(centroids, np.float32).reshape(n, 3),
                  need_viewdir=np.array(bool(need_viewdir)),
                  need_appearance_embedding=np.array(bool(need_appearance_embedding)))
    f = lambda a: np.ascontiguousarray(a, np.float32)
    for i, sm in enumerate(submodules):
        p = f"sub_module_{i}/"
        nt = len(sm["trunk_w"])
        arrays[p + "config"] = np.asarray([nt, sm.get("skip_layer", 4), sm.get("pe_xyz_freqs", 12),
                                           sm.get("pe_dir_freqs", 4), sm.get("sigma_activation", 1)], np.int32)
        for l in range(nt):
            arrays[p + f"trunk_w_{l}"] = f(sm["trunk_w"][l])
            arrays[p + f"trunk_b_{l}"] = f(sm["trunk_b"][l])
        for k in ("sigma_w", "sigma_b", "final_w", "final_b", "head1_w", "head1_b", "head2_w", "head2_b"):
            arrays[p + k] = f(sm[k])
        if sm.get("embedding") is not None:
            arrays[p + "embedding"] = f(sm["embedding"])
    (np.savez_compressed if compressed else np.savez)(path, **arrays)


HEADLESS_BIN = os.path.join(_HERE, "bin", "mnv_headless")


def bytes_checksum(a: np.ndarray) -> str:
    """The 64-bit position-weighted byte sum `mnv_headless --selftest-*` prints."""
    b = np.ascontiguousarray(a).view(np.uint8).ravel().astype(np.uint64)
    i = np.arange(b.size, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return f"{int(((b + np.uint64(1)) * (i * np.uint64(2654435761) + np.uint64(1))).sum(dtype=np.uint64)):016x}"
