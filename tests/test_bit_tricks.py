"""Integer identities the march kernels rely on (csrc/mnv_render.cu, csrc/mnv_march.cuh), restated in numpy: a wrong
constant here is a wrong leaf on the GPU, and these run without one."""
import numpy as np


def _funnel_l(lo, hi, n):
    """CUDA __funnelshift_l(lo, hi, n) for 0 < n < 32: upper word of (hi:lo) << n."""
    lo, hi = lo.astype(np.uint64), hi.astype(np.uint64)
    return ((((hi << np.uint64(32)) | lo) << np.uint64(n)) >> np.uint64(32)).astype(np.uint64) & np.uint64(0xFFFFFFFF)


def test_child_slot_by_chained_funnel_shifts():
    """slot = node * 8 + child(level) from q = 0x4B000000 | 23-bit cell coordinate per axis: the three funnel shifts of
    render_pixel / march_step against the shift-mask-or form they replaced, for every level and descent round."""
    rng = np.random.default_rng(0)
    n = 20000
    q = [(rng.integers(0, 1 << 23, n, dtype=np.uint64) | np.uint64(0x4B000000)) for _ in range(3)]
    node = rng.integers(0, 1 << 28, n, dtype=np.uint64)
    for lvl in range(0, 23):
        s = [(v << np.uint64(9 + lvl)) & np.uint64(0xFFFFFFFF) for v in q]  # level bit of each axis at bit 31
        for rnd in range(0, 23 - lvl):
            bit = 22 - (lvl + rnd)
            child = (((q[0] >> np.uint64(bit)) & np.uint64(1)) << np.uint64(2)) | (((q[1] >> np.uint64(bit)) & np.uint64(1)) << np.uint64(1)) | \
                    ((q[2] >> np.uint64(bit)) & np.uint64(1))
            want = (node * np.uint64(8) + child) & np.uint64(0xFFFFFFFF)
            got = _funnel_l(s[2], _funnel_l(s[1], _funnel_l(s[0], node, 1), 1), 1)
            assert np.array_equal(got, want), (lvl, rnd)
            s = [(v << np.uint64(1)) & np.uint64(0xFFFFFFFF) for v in s]


def test_cell_coordinate_from_the_magic_constant():
    """q = bits(fma_rd(p, 2^23, 2^23)) carries floor(p * 2^23) in its mantissa for p in [0, 1)."""
    rng = np.random.default_rng(1)
    p = rng.random(100000).astype(np.float32)
    p = np.minimum(p, np.float32(1.0) - np.float32(1e-6))
    exact = np.floor(p.astype(np.float64) * 8388608.0)  # exact in double
    q = (exact + 8388608.0).astype(np.float32).view(np.uint32)  # the value fma_rd produces (representable exactly)
    assert np.array_equal(q & np.uint32(0x7FFFFF), exact.astype(np.uint32))
    assert np.all(q >> np.uint32(23) == np.uint32(0x4B000000 >> 23))


def test_common_ancestor_level_from_xor():
    """lvl = clz(diff) - 9 with diff = OR of the per-axis XORs of consecutive q: the number of leading cell-coordinate
    bits two positions share = the deepest level whose node contains both."""
    rng = np.random.default_rng(2)
    a = rng.integers(0, 1 << 23, (5000, 3), dtype=np.uint64)
    flip = rng.integers(0, 23, 5000)
    b = a.copy()
    axis = rng.integers(0, 3, 5000)
    b[np.arange(5000), axis] ^= (np.uint64(1) << flip.astype(np.uint64))           # first differing bit = `flip`
    b[np.arange(5000), axis] ^= rng.integers(0, 1 << 23, 5000, dtype=np.uint64) & ((np.uint64(1) << flip.astype(np.uint64)) - np.uint64(1))
    diff = np.bitwise_or.reduce((a | np.uint64(0x4B000000)) ^ (b | np.uint64(0x4B000000)), axis=1)
    clz = 32 - np.floor(np.log2(diff.astype(np.float64))).astype(int) - 1
    assert np.array_equal(clz - 9, 22 - flip)  # levels 0 .. 22-flip-1 are shared: descent restarts at level 22 - flip
