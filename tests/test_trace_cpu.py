"""The teacher-forcing trace (oracle/ref_harness.cpp: pmref_trace_dest) walks Propagate::propagatePatch's control flow
(pmmvps/propagate.cpp:126-218) through the reference's own member functions so that every try's intermediate patch can be
recorded.  This test pins that walk to the REAL propagatePatch: from the same store state both must leave bit-identical stores,
dest cell after dest cell, in the fill branch (m_depth 1) and in the challenge branch with Optim::check (m_depth 2).  CPU only."""
import os

import numpy as np
import pytest

SEED = 0x5EED0001C0FFEE


@pytest.fixture(scope="module")
def tiny(tmp_path_factory):
    from mvskit_b200 import synth
    from oracle import pyoracle
    if not os.path.exists(pyoracle.REF_SO):
        if os.path.isdir("/root/reference/pmmvps"):
            pyoracle.build(ref=True)
        else:
            pytest.skip("oracle/_ref/libpmref.so not built and /root/reference absent")
    scene = synth.make_scene(1, scale=0.4).render()
    d = synth.write_scene(scene, str(tmp_path_factory.mktemp("scene_trace")))
    return scene, pyoracle.RefLib(d)


def _restore(ref, pb, depth):
    ref.clear_patches()
    ref.set_depth(0)
    ref.add_patches(pb.coord, pb.normal, pb.scal, pb.images, pb.nimages)
    ref.set_depth(depth)
    if depth >= 1:
        ref.filter_rebuild(0)
    ref.refine_seed(SEED)


def _same(a, b):
    assert a.n == b.n, (a.n, b.n)
    for name in ("coord", "normal", "scal"):
        x, y = getattr(a, name), getattr(b, name)
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), name
    for name in ("images", "nimages", "grids", "vimages", "nvimages", "vgrids"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name


def test_traced_walk_equals_the_real_propagate_patch(tiny):
    scene, ref = tiny
    ref.clear_patches(); ref.set_depth(0); ref.set_ncc_thresholds(0.7, 0.4); ref.create_patches()
    state = ref.get_patches()
    img = 0
    gw, gh = ref.grid_dims(img)
    codes = np.zeros(4, int)
    decisions = np.zeros(3, int)
    # ---- fill branch: grow the store diagonal by diagonal, comparing the two walks on every 3rd diagonal ----
    _restore(ref, state, 1)
    for d in range(8, 75):
        if d % 3 == 0:
            before = ref.get_patches()
            _restore(ref, before, 1)                         # both walks start from the same re-loaded records
            calls_a = ref.propagate_diag(img, d, 1, 0)
            a = ref.get_patches()
            _restore(ref, before, 1)
            calls_b = 0
            for x in range(max(0, d - gh + 1), min(gw - 1, d) + 1):
                tr = ref.trace_dest(img, x, d - x, 1, 0)
                calls_b += tr.n // 2
                for t in range(tr.n):
                    codes[tr.code[t]] += 1
                    decisions[tr.decision[t]] += 1
            b = ref.get_patches()
            assert calls_a == calls_b
            _same(a, b)
        else:
            ref.propagate_diag(img, d, 1, 0)
    assert codes[3] > 50 and decisions[1] > 30, (codes, decisions)
    # ---- challenge branch with Optim::check: the cells are full now, m_depth = 2, reverse sweep ----
    for v in (1, 2, 3):                                      # more views register their patches in view 0's cells until they are full
        for d in range(gw + gh - 1):
            ref.propagate_diag(v, d, 1, 0)
    grown = ref.get_patches()
    ndiag = gw + gh - 1
    full = np.zeros(2, int)
    for k in range(ndiag - 1 - 60, ndiag - 1 - 30, 3):
        _restore(ref, grown, 2)
        calls_a = ref.propagate_diag(img, k, -1, 1)
        a = ref.get_patches()
        _restore(ref, grown, 2)
        for x in range(max(0, k - gh + 1), min(gw - 1, k) + 1):
            tr = ref.trace_dest(img, x, k - x, -1, 1)
            for t in range(tr.n):
                codes[tr.code[t]] += 1
                decisions[tr.decision[t]] += 1
                full[tr.branch_full[t]] += 1
        b = ref.get_patches()
        _same(a, b)
    print("trace walk:", dict(codes=codes.tolist(), decisions=decisions.tolist(), full=full.tolist()))
    assert full[1] > 20 and codes[1] > 5, (full, codes)          # the challenge branch ran, and challengers lost
