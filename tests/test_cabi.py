"""CPU tests of the drop-in boundary: libmnv_b200.so loads, exports every symbol
include/mnv_b200.h declares, and fails loudly (never falls back) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "mnv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mnv_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    names = header_functions()
    for must in ("mnv_tree_create", "mnv_render_voxels", "mnv_render_frame_host", "mnv_query_points"):
        assert must in names


def test_library_exports_every_declared_symbol(mnv):
    L = mnv.lib()
    missing = [n for n in header_functions() if not hasattr(L, n)]
    assert not missing, f"declared in include/mnv_b200.h but not exported: {missing}"


def test_render_options_layout_and_defaults(mnv):
    o = mnv.default_options()
    # include/render_options.hpp:12-55
    assert o.step_size == pytest.approx(1e-4) and o.sigma_thresh == pytest.approx(1e-2)
    assert o.stop_thresh == pytest.approx(1e-2) and o.background_brightness == 1.0
    assert list(o.render_bbox) == [0, 0, 0, 1, 1, 1] and list(o.basis_minmax) == [0, 24]
    assert (o.max_depth, o.samples_per_corner, o.split_batch_size, o.nerf_batch_size) == (16, 8, 4192, 1024)
    assert (o.max_sample_count, o.appearance_embedding, o.max_guided_samples) == (256, -1, 128)
    assert not o.use_splitting and not o.use_guided_sampling and not o.need_viewdir
    assert C.sizeof(o) == 104  # sizeof(viewer::RenderOptions) on this ABI


def test_no_cpu_fallback(mnv):
    """Without a device the product path raises; it never routes to the oracle."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    tree = mnv.synth.make_tree(depth=3)
    with pytest.raises(mnv.MnvError) as ei:
        mnv.DeviceTree(tree)
    assert ei.value.code == 2  # MNV_ERR_NO_DEVICE


def test_invalid_trees_are_rejected(mnv):
    tree = mnv.synth.make_tree(depth=3)
    d = mnv.TreeDesc()
    d.N, d.data_dim, d.format, d.basis_dim, d.capacity = 3, 28, 1, 9, tree.capacity
    data = np.ascontiguousarray(tree.data.view(np.uint16))
    child = np.ascontiguousarray(tree.child)
    d.data, d.child = data.ctypes.data, child.ctypes.data
    h = C.c_void_p()
    assert mnv.lib().mnv_tree_create(C.byref(h), C.byref(d), 0, 0) == 1  # N != 2
    assert b"N == 2" in mnv.lib().mnv_last_error()
    d.N, d.basis_dim = 2, 7
    assert mnv.lib().mnv_tree_create(C.byref(h), C.byref(d), 0, 0) == 6  # MNV_ERR_FORMAT


def test_tracker_chunk_encoding_roundtrip(mnv):
    """Chunk column of the candidate trackers: the reference's float value below 2^24 (bit-identical rows),
    integer bits above (rt_core.cuh:238-240 loses odd ids there) — host helpers of the C-ABI."""
    L = mnv.lib()
    vals = [0, 1, 7, 12345, (1 << 24) - 1, 1 << 24, (1 << 24) + 1, 18_200_001, (1 << 27) + 12345, (1 << 28) - 1]
    for v in vals:
        f = L.mnv_tracker_encode_chunk(v)
        assert L.mnv_tracker_decode_chunk(f) == v
        if v < (1 << 24):
            assert f == float(v)  # exactly what the reference writes
        else:
            assert 0.0 < f < 1e-28  # cannot be mistaken for a float-encoded id, 0 or -1
    col = np.array([L.mnv_tracker_encode_chunk(v) for v in vals] + [-1.0], np.float32)
    assert np.array_equal(mnv.decode_tracker_chunks(col), np.array(vals + [-1]))
