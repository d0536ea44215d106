"""GPU parity (through the C ABI) of K0 (pyramid) and K1 (batched hypothesis NCC) against the oracle.

Bar (BASELINE.json north_star): integer work bit-exact (cell indices, pyramid level chosen per view, valid /
invalid decisions), NCC scores within 1e-4 absolute.  Everything that feeds an integer decision is also
checked bit-exact in float (projection, getUnit, patch axes), because that is how the decisions stay exact.
"""
import numpy as np
import pytest

from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu

NCC_TOL = 1e-4     # absolute, north_star


@pytest.fixture(scope="module")
def ctx(small_scene):
    from mvskit_b200 import pmk
    c = pmk.Context(nviews=small_scene.nviews)
    c.set_scene(small_scene.P, small_scene.images)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx_sphere(sphere_scene):
    from mvskit_b200 import pmk
    c = pmk.Context(nviews=sphere_scene.nviews)
    c.set_scene(sphere_scene.P, sphere_scene.images)
    yield c
    c.close()


def test_camera_constants_bit_exact(ctx, coracle, small_scene):
    for v in range(small_scene.nviews):
        for lvl in range(ctx.nlevels):
            a, b = ctx.camera(v, lvl), coracle.camera(v, lvl)
            for k in b:
                assert_bits_equal(np.asarray(a[k]), np.asarray(b[k]), f"camera[{v}].{k}@{lvl}")


def test_pyramid_bit_exact(ctx, coracle, small_scene):
    for v in range(small_scene.nviews):
        for lvl in range(ctx.nlevels):
            assert ctx.level_dims(v, lvl) == coracle.image_dims(v, lvl)
            assert np.array_equal(ctx.level_image(v, lvl), coracle.image(v, lvl)), (v, lvl)
        gw, gh = ctx.grid_dims(v)
        w, h = coracle.image_dims(v, 1)
        assert (gw, gh) == ((w + 1) // 2, (h + 1) // 2)


def test_building_blocks_bit_exact(ctx, coracle, hyps):
    c, n, vw, nv = hyps
    v0 = vw[:, 0].copy()
    got = ctx.probe(v0, c, n)
    assert_bits_equal(got["project"], coracle.project(v0, c), "project")
    assert_bits_equal(got["unit"], coracle.get_unit(v0, c), "getUnit")
    px, py = coracle.get_paxes(v0, c, n)
    assert_bits_equal(got["px"], px, "pxaxis")
    assert_bits_equal(got["py"], py, "pyaxis")
    ixy, ok = coracle.cells(v0, c)
    assert np.array_equal(got["cell"], ixy)
    assert np.array_equal(got["cell_ok"], ok)


def test_unproject_bit_exact(ctx, coracle, reflib, hyps, small_scene):
    """Camera::unproject (camera.cpp:329-337) as Propagate::generatePatch drives it (propagate.cpp:224-226): depth * (u, v, 1) of a
    pixel of the reference view back to 3-D -- against the C restatement AND against the compiled reference itself."""
    c, n, vw, nv = hyps
    v0 = vw[:, 0].copy()
    rng = np.random.RandomState(11)
    depth = rng.uniform(1.0, 6.0, len(v0)).astype(np.float32)
    uv = np.stack([rng.uniform(-20, small_scene.width // 2 + 20, len(v0)), rng.uniform(-20, small_scene.height // 2 + 20, len(v0)), np.ones(len(v0))], 1).astype(np.float32)
    ic = (uv * depth[:, None]).astype(np.float32)                     # float32 products, as `depth * icoord` in the reference
    got = ctx.probe_unproject(v0, ic)
    assert_bits_equal(got, coracle.unproject(v0, ic), "unproject vs the C restatement")
    assert_bits_equal(got, reflib.unproject(v0, ic), "unproject vs libpmref.so")
    # round trip: the projection of the unprojected point is the pixel again (float32 tolerance)
    back = ctx.probe(v0, got)["project"]
    assert np.abs(back[:, :2] - uv[:, :2]).max() < 2e-2


def test_cells_behind_camera_and_negative(ctx, coracle, small_scene):
    # points behind the camera hit the (-65535,-65535,-1) sentinel; points just left of the image exercise the
    # truncating integer division quirk (patch_manager.cpp:230-233): x in (-csize-0.5, -0.5) -> cell 0
    rng = np.random.RandomState(3)
    c = np.concatenate([rng.uniform(-6, 6, (2000, 3)), np.ones((2000, 1))], 1).astype(np.float32)
    v = rng.randint(0, small_scene.nviews, 2000).astype(np.int32)
    got = ctx.probe(v, c)
    ixy, ok = coracle.cells(v, c)
    assert_bits_equal(got["project"], coracle.project(v, c), "project")
    assert np.array_equal(got["cell"], ixy) and np.array_equal(got["cell_ok"], ok)
    assert (got["project"][:, 2] == -1).any()


def _check_ncc(ctx, oracle, hyp):
    c, n, vw, nv = hyp
    incc, ncc, lv = ctx.ncc_eval(c, n, vw, nv, want_levels=True)
    oi, on, ol = oracle.compute_ncc(c, n, vw, nv, True)
    assert np.array_equal(lv, ol), "pyramid level / validity per view must be bit-exact"
    assert np.array_equal(incc == 2.0, oi == 2.0), "invalid-score sentinel decisions must match"
    good = oi != 2.0
    assert np.abs(incc[good] - oi[good]).max() <= NCC_TOL
    assert np.abs(ncc[good] - on[good]).max() <= 4 * NCC_TOL     # ncc = 1 - x/(1-3x): slope <= 4 for incc <= 1/6
    return incc, oi


def test_ncc_parity_plane(ctx, coracle, hyps):
    incc, oi = _check_ncc(ctx, coracle, hyps)
    assert (oi < 0.3).mean() > 0.5


def test_ncc_parity_sphere_with_occlusion(ctx_sphere, coracle_sphere, sphere_scene):
    hyp = sphere_scene.hypotheses(4096, seed=11, depth_jitter=0.03, normal_jitter_deg=30.0, well_observed=False)
    _check_ncc(ctx_sphere, coracle_sphere, hyp)


def test_ncc_vs_reference_itself(ctx, reflib, hyps):
    c, n, vw, nv = hyps
    incc, ncc = ctx.ncc_eval(c, n, vw, nv)
    ri, rn = reflib.compute_ncc(c, n, vw, nv)
    assert np.array_equal(incc == 2.0, ri == 2.0)
    good = ri != 2.0
    assert np.abs(incc[good] - ri[good]).max() <= NCC_TOL


def test_ncc_edge_cases(ctx, coracle, small_scene, hyps):
    c, n, vw, nv = (a.copy() for a in hyps)
    m = 64
    c, n, vw, nv = c[:m], n[:m], vw[:m], nv[:m]
    nv[0] = 0                      # no images
    nv[1] = 1                      # only the reference image -> 2.0 (optim.cpp:631)
    n[2] = -n[2]                   # back-facing: every view fails the 60 degree gate
    c[3, :3] += 50.0               # far outside every image
    vw[4, 1] = vw[4, 0]            # duplicate view
    c[5, :3] = small_scene.eyes[0] + 1e-3   # almost at a camera centre
    n[6, :3] = 0.0                 # zero normal -> NaNs in getPAxes; must not crash, decisions must match
    incc, ncc, lv = ctx.ncc_eval(c, n, vw, nv, want_levels=True)
    oi, on, ol = coracle.compute_ncc(c, n, vw, nv, True)
    assert np.array_equal(lv, ol)
    both = np.isfinite(oi)
    assert np.array_equal(np.isfinite(incc), both)
    assert np.abs(incc[both] - oi[both]).max() <= NCC_TOL
    assert incc[0] == 2.0 and incc[1] == 2.0 and incc[2] == 2.0 and incc[3] == 2.0
    # empty batch and ragged tails
    e = ctx.ncc_eval(c[:0], n[:0], vw[:0], nv[:0])
    assert e[0].shape == (0,)
    for k in (1, 31, 33):
        a = ctx.ncc_eval(c[:k], n[:k], vw[:k], nv[:k])[0]
        assert np.array_equal(a, incc[:k], equal_nan=True)


def test_ncc_batch_invariance_large(ctx, small_scene):
    """Size-independent properties at a size the oracle is not run on: results do not depend on the order
    or the batching of hypotheses (each eval is independent), and repeated launches are deterministic."""
    c, n, vw, nv = small_scene.hypotheses(1 << 18, seed=99, well_observed=False)
    a = ctx.ncc_eval(c, n, vw, nv)[0]
    b = ctx.ncc_eval(c, n, vw, nv)[0]
    assert_bits_equal(a, b, "determinism")
    perm = np.random.RandomState(0).permutation(len(c))
    p = ctx.ncc_eval(c[perm], n[perm], vw[perm], nv[perm])[0]
    assert_bits_equal(p, a[perm], "permutation invariance")
    h = len(c) // 3
    s = np.concatenate([ctx.ncc_eval(c[:h], n[:h], vw[:h], nv[:h])[0], ctx.ncc_eval(c[h:], n[h:], vw[h:], nv[h:])[0]])
    assert_bits_equal(s, a, "split invariance")
    ok = a != 2.0
    assert ok.mean() > 0.5 and (a[ok] >= 0).all() and (a[ok] < 0.7).all()     # robust incc range: x/(1+3x) < 1/3 .. 2/7


def test_streamed_host_call_equals_the_device_call(ctx, small_scene):
    """pmk_ncc_eval with host buffers streams the hypotheses over PCIe into ONE running K1 launch (per-chunk arrival words) and,
    with mapped pinned outputs, lets the kernel write the scores straight into the caller's buffers.  Whatever the buffers are
    (pinned or pageable) and wherever the chunk boundaries fall (ragged tail, fewer hypotheses than one chunk), the results are the
    bits of the device-pointer call on the same inputs."""
    from mvskit_b200 import pmk
    N = 3 * (1 << 17) + 12345
    c, n, vw, nv = small_scene.hypotheses(N, seed=123, well_observed=False)
    d = [ctx.alloc(a.nbytes).upload(a) for a in (c, n, vw, nv)]
    d_incc, d_ncc, d_lv = ctx.alloc(N * 4), ctx.alloc(N * 4), ctx.alloc(N * ctx.tau * 4)
    ctx.ncc_eval_dev(N, d[0], d[1], d[2], d[3], vw.shape[1], d_incc, d_ncc, d_lv)
    want_incc, want_ncc, want_lv = np.empty(N, np.float32), np.empty(N, np.float32), np.empty((N, ctx.tau), np.int32)
    d_incc.download(want_incc); d_ncc.download(want_ncc); d_lv.download(want_lv)
    # pageable numpy buffers
    incc, ncc, lv = ctx.ncc_eval(c, n, vw, nv, want_levels=True)
    assert_bits_equal(incc, want_incc, "pageable incc"); assert_bits_equal(ncc, want_ncc, "pageable ncc"); assert np.array_equal(lv, want_lv)
    # pinned buffers: zero-copy outputs; run twice so stale arrival words of the previous call are in place
    h = [pmk.pinned_empty(a.shape, a.dtype) for a in (c, n, vw, nv)]
    for dst, src in zip(h, (c, n, vw, nv)):
        dst[:] = src
    h_incc, h_ncc = pmk.pinned_empty((N,), np.float32), pmk.pinned_empty((N,), np.float32)
    for m in (N, 1000, N - 7, 1, 31, (1 << 14) - 1, 1 << 14, (1 << 14) + 1, (1 << 17) + 33):     # slot (2^14) and chunk (2^17) boundaries
        h_incc[:] = -7.0; h_ncc[:] = -7.0
        pmk._chk(pmk.lib().pmk_ncc_eval(ctx.h, m, pmk._p(h[0]), pmk._p(h[1]), pmk._p(h[2]), pmk._p(h[3]), vw.shape[1], pmk._p(h_incc), pmk._p(h_ncc), None))
        assert_bits_equal(h_incc[:m], want_incc[:m], f"pinned incc n={m}"); assert_bits_equal(h_ncc[:m], want_ncc[:m], f"pinned ncc n={m}")
        assert (h_incc[m:] == -7.0).all()
    for b in d + [d_incc, d_ncc, d_lv]:
        b.free()


def test_packed_host_call_is_bit_identical(ctx, small_scene):
    """pmk_ncc_eval_packed (3-float coordinates / normals, byte view ids: 31 B per hypothesis instead of 60) against pmk_ncc_eval on the
    widened records -- same bits, scores and pyramid levels, through pageable and pinned buffers, ragged sizes."""
    from mvskit_b200 import pmk
    N = 2 * (1 << 17) + 777
    c, n, vw, nv = small_scene.hypotheses(N, seed=321, well_observed=False)
    assert (c[:, 3] == 1).all()
    n = n.copy(); n[:, 3] = 0.0
    incc, ncc, lv = ctx.ncc_eval(c, n, vw, nv, want_levels=True)
    c3, n3, v8, nv8 = ctx.pack_hypotheses(c, n, vw, nv)
    pi, pn, pl = ctx.ncc_eval_packed(c3, n3, v8, nv8, want_levels=True)
    assert_bits_equal(pi, incc, "packed incc"); assert_bits_equal(pn, ncc, "packed ncc"); assert np.array_equal(pl, lv)
    h = [pmk.pinned_empty(a.shape, a.dtype) for a in (c3, n3, v8, nv8)]
    for dst, src in zip(h, (c3, n3, v8, nv8)):
        dst[:] = src
    out = (pmk.pinned_empty((N,), np.float32), pmk.pinned_empty((N,), np.float32))
    for m in (N, 1, 31, (1 << 14) + 1, (1 << 17) - 5):
        out[0][:] = -7.0
        ctx.ncc_eval_packed(h[0][:m], h[1][:m], h[2][:m], h[3][:m], out=(out[0][:m], out[1][:m]))
        assert_bits_equal(out[0][:m], incc[:m], f"pinned packed incc n={m}"); assert_bits_equal(out[1][:m], ncc[:m], f"pinned packed ncc n={m}")
        assert (out[0][m:] == -7.0).all()


def test_state_errors(small_scene):
    from mvskit_b200 import pmk
    c = pmk.Context(nviews=small_scene.nviews)
    c.set_view(0, small_scene.P[0], small_scene.images[0])
    co, no, vw, nv = small_scene.hypotheses(8)
    with pytest.raises(pmk.PmkError, match="has not been uploaded"):
        c.ncc_eval(co, no, vw, nv)
    with pytest.raises(pmk.PmkError, match="already uploaded"):
        c.set_view(0, small_scene.P[0], small_scene.images[0])
    c.close()
