"""Property tests of the C++ .npz reader / writer behind viewer::N3Tree (csrc/viewer/npz.cpp): whatever tree shape,
data format and compression numpy writes, the loader sees the same bytes; whatever the C++ writer emits, numpy and
zipfile read back.  Exercised through the `mnv_headless --selftest-*` modes (no GPU)."""
import json
import subprocess
import zipfile

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st


def _run(mnv, *args):
    return subprocess.run([mnv.HEADLESS_BIN, *map(str, args)], capture_output=True, text=True, timeout=120)


def _random_tree(rng, n_internal, data_dim, fmt, mnv):
    """A random valid N3Tree: node i > 0 hangs under a random leaf slot of an earlier node."""
    cap = n_internal
    child = np.zeros((cap, 8), np.int32)
    parent = np.zeros(cap, np.int32)
    depth = np.zeros(cap, np.int32)
    for i in range(1, cap):
        while True:
            pn, pc = int(rng.integers(0, i)), int(rng.integers(0, 8))
            if child[pn, pc] == 0:
                break
        child[pn, pc] = i - pn
        parent[i] = pn * 8 + pc
        depth[i] = depth[pn] + 1
    data = rng.standard_normal((cap, 8, data_dim)).astype(np.float16)
    return mnv.HostTree(N=2, data_dim=data_dim, data_format=fmt, child=child, parent=parent, depth=depth, data=data,
                        scale=rng.uniform(0.1, 2.0, 3).astype(np.float32), offset=rng.uniform(-1, 1, 3).astype(np.float32))


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 40), fmt=st.sampled_from(["RGBA", "SH1", "SH4", "SH9", "SH16", "SH25"]),
       compressed=st.booleans())
def test_reader_and_writer_round_trip_random_trees(mnv, tmp_path_factory, seed, n, fmt, compressed):
    mnv.build_library()
    d = tmp_path_factory.mktemp("npz")
    rng = np.random.default_rng(seed)
    data_dim = 4 if fmt == "RGBA" else 3 * int(fmt[2:]) + 1
    tree = _random_tree(rng, n, data_dim, fmt, mnv)
    src, dst = d / "a.npz", d / "b.npz"
    tree.save_npz(str(src), compressed=compressed)
    r = _run(mnv, src, "--selftest-load")
    assert r.returncode == 0, r.stderr
    j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert j["capacity"] == n and j["data_dim"] == data_dim and j["format"] == fmt
    assert j["child"] == mnv.bytes_checksum(tree.child) and j["parent"] == mnv.bytes_checksum(tree.parent)
    assert j["data"] == mnv.bytes_checksum(tree.data.view(np.uint16))
    assert np.allclose(j["scale"], tree.scale, rtol=1e-7) and np.allclose(j["offset"], tree.offset, rtol=1e-7)
    # writer: C++ -> numpy
    r = _run(mnv, src, "--selftest-resave", dst)
    assert r.returncode == 0, r.stderr
    with zipfile.ZipFile(dst) as z:
        assert z.testzip() is None
    back = mnv.HostTree.load_npz(str(dst))
    assert back.data_format == fmt and np.array_equal(back.child, tree.child) and np.array_equal(back.parent, tree.parent)
    assert np.array_equal(back.depth, tree.depth) and np.array_equal(back.data.view(np.uint16), tree.data.view(np.uint16))


@settings(max_examples=20, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(cut=st.floats(0.02, 0.98), flip=st.integers(0, 10 ** 6))
def test_damaged_archives_fail_cleanly(mnv, tmp_path_factory, cut, flip):
    """Truncated or bit-flipped files must produce an error exit (or load if the damage missed everything that is
    parsed) — never a crash / hang."""
    mnv.build_library()
    d = tmp_path_factory.mktemp("bad")
    tree = mnv.synth.make_tree(depth=3)
    src = d / "a.npz"
    tree.save_npz(str(src), compressed=True)
    blob = bytearray(src.read_bytes())
    (d / "trunc.npz").write_bytes(bytes(blob[: max(1, int(len(blob) * cut))]))
    r = _run(mnv, d / "trunc.npz", "--selftest-load")
    assert r.returncode in (0, 1), (r.returncode, r.stderr[-300:])
    blob[flip % len(blob)] ^= 0x5A
    (d / "flip.npz").write_bytes(bytes(blob))
    r = _run(mnv, d / "flip.npz", "--selftest-load")
    assert r.returncode in (0, 1), (r.returncode, r.stderr[-300:])
