"""The carry rule of guided sampling with sharded sub-modules (csrc/mnv_guided.cu `segment_context`, mirrored on the
host by multigpu.segment_context_host): splitting a ray's leaf sequence into per-cell segments, probing each segment
alone and carrying (T, sample count) across the segments in first-z order must select exactly the samples the
unsharded emission loop (rt_core.cuh:321-357: emit while count < max_guided_samples, T *= att, stop once
T < stop_thresh) selects, and name the right 'next segment' for the compositor.  Attenuations are powers of two so
the products are exact in any order."""
import os
import sys

import numpy as np
from hypothesis import given, settings, strategies as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _march(leaves, T, count, thresh, cap):
    """leaves: [(id, att)] -> (emitted ids, T at the end, samples emitted); the reference's loop."""
    out, n = [], 0
    if T < thresh or count >= cap:
        return out, T, n
    for lid, att in leaves:
        if count + n < cap:
            out.append(lid)
            n += 1
        T *= att
        if T < thresh:
            break
    return out, T, n


@st.composite
def rays(draw):
    n_cells = draw(st.integers(1, 8))
    order = draw(st.permutations(list(range(n_cells))))
    segs, lid, z = {}, 0, 1.0
    for c in order:
        k = draw(st.integers(0, 6))
        seg = []
        for _ in range(k):
            seg.append((lid, 2.0 ** -draw(st.integers(0, 4)), z))
            lid += 1
            z += 0.25
        segs[c] = seg
    thresh = 2.0 ** -draw(st.integers(1, 12)) * 1.5
    cap = draw(st.integers(1, 24))
    return n_cells, order, segs, thresh, cap


@settings(max_examples=400, deadline=None)
@given(rays())
def test_carried_segments_select_the_unsharded_samples(ray):
    import mega_nerf_viewer_b200.multigpu as MG

    n_cells, order, segs, thresh, cap = ray
    whole = [(lid, att) for c in order for lid, att, _ in segs[c]]
    z_of = {lid: z for c in order for lid, _, z in segs[c]}
    want, _, _ = _march(whole, 1.0, 0, thresh, cap)
    # probe every cell alone
    records = []
    for c in range(n_cells):
        em, T, n = _march([(l, a) for l, a, _ in segs[c]], 1.0, 0, thresh, cap)
        records.append((T, n, z_of[em[0]] if em else MG.NO_SEGMENT))
    got, firsts, nexts = [], {}, {}
    for c in range(n_cells):
        T_in, count_in, z_next = MG.segment_context_host(records, c, thresh, cap)
        em, _, _ = _march([(l, a) for l, a, _ in segs[c]], T_in, count_in, thresh, cap)
        got += em
        if em:
            firsts[c] = z_of[em[0]]
            nexts[c] = z_next
    assert sorted(got) == want
    # the compositor's 'next segment': the first sample of the next cell that emitted, none for the last one
    emitting = sorted(firsts, key=lambda c: firsts[c])
    for a, b in zip(emitting, emitting[1:] + [None]):
        assert nexts[a] == (firsts[b] if b is not None else MG.NO_SEGMENT)


def test_carry_examples():
    import mega_nerf_viewer_b200.multigpu as MG

    # two cells, the front one opaque: the back one must not emit
    rec = [(0.001, 3, 1.0), (0.5, 2, 2.0)]
    assert MG.segment_context_host(rec, 1, 0.01, 64)[0] < 0.01
    assert MG.segment_context_host(rec, 0, 0.01, 64) == (1.0, 0, MG.NO_SEGMENT)
    # cap reached in front: the back cell starts with the full count; a cell without samples is skipped
    rec = [(0.9, 4, 1.0), (1.0, 0, MG.NO_SEGMENT), (0.9, 4, 3.0)]
    assert MG.segment_context_host(rec, 2, 0.01, 4)[1] == 4
    assert MG.segment_context_host(rec, 0, 0.01, 4)[2] == MG.NO_SEGMENT
    assert MG.segment_context_host(rec, 0, 0.01, 5)[2] == 3.0
