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


def _load_rc(mnv, path):
    import subprocess

    r = subprocess.run([mnv.HEADLESS_BIN, str(path), "--selftest-load"], capture_output=True, timeout=30)
    return r.returncode, r.stderr[-200:].decode("latin1")


def test_npy_header_damage_behind_a_valid_crc(mnv, tmp_path):
    """Members re-zipped with correct CRCs but damaged .npy headers, truncated payloads or absurd header lengths:
    the loader must reject (or load) them, never crash — the parser may not lean on the CRC check."""
    import random
    import zipfile

    mnv.build_library()
    rnd = random.Random(7)
    tree = mnv.synth.make_tree(depth=3)
    src = tmp_path / "a.npz"
    tree.save_npz(str(src), compressed=False)
    with zipfile.ZipFile(src) as z:
        members = {n: z.read(n) for n in z.namelist()}
    alphabet = b"0123456789(),' <>|fiuUS:{}x\x00\xff"
    for k in range(150):
        m = dict(members)
        name = rnd.choice(sorted(m))
        b = bytearray(m[name])
        mode = rnd.random()
        if mode < 0.6:
            hl = b[8] | (b[9] << 8)
            for _ in range(rnd.choice((1, 2, 3))):
                b[10 + rnd.randrange(hl)] = rnd.choice(alphabet)
        elif mode < 0.8:
            b = b[: rnd.randrange(6, len(b))]
        else:
            b[8], b[9] = rnd.randrange(256), rnd.randrange(256)
        m[name] = bytes(b)
        dst = tmp_path / "y.npz"
        with zipfile.ZipFile(dst, "w", zipfile.ZIP_DEFLATED if k % 2 else zipfile.ZIP_STORED) as z:
            for n, v in m.items():
                z.writestr(n, v)
        rc, err = _load_rc(mnv, dst)
        assert rc in (0, 1), (k, name, rc, err, bytes(b[:100]))


def test_well_formed_archives_with_hostile_contents(mnv, tmp_path):
    """Arrays numpy wrote correctly but that do not describe a tree: every one must be refused with a message
    (reference: std::runtime_error on bad dtype / shape, src/n3tree/n3tree.cpp:114,121,180,196,200)."""
    mnv.build_library()
    tree = mnv.synth.make_tree(depth=3)
    src = tmp_path / "a.npz"
    tree.save_npz(str(src))
    base = dict(np.load(src))

    def variant(**kw):
        c = dict(base)
        c.update(kw)
        return c

    far, back = base["child"].copy(), base["child"].copy()
    far[0, 0, 0, 0] = 10 ** 6
    back[-1, 1, 1, 1] = -(10 ** 6)
    cases = {
        "child link out of range": [variant(child=far), variant(child=back)],
        "data_dim": [variant(data_dim=np.int64(10 ** 6)), variant(data_dim=np.int64(-1))],
        "child must be": [variant(child=base["child"][:, :, :, :1]), variant(child=np.zeros((0, 2, 2, 2), np.int32))],
        "not aligned": [variant(parent_depth=base["parent_depth"][:3])],
        "data shape": [variant(data=base["data"][..., :5])],
        "invradius3": [variant(invradius3=np.zeros(1, np.float32))],
        "offset": [variant(offset=np.zeros(3, np.float64))],
        "half precision": [variant(data=base["data"].astype(np.float32))],
    }
    for msg, variants in cases.items():
        for i, arrays in enumerate(variants):
            dst = tmp_path / "z.npz"
            np.savez(dst, **arrays)
            rc, err = _load_rc(mnv, dst)
            assert rc == 1 and msg in err, (msg, i, rc, err)
