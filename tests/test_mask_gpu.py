"""GPU parity of the silhouette-mask path (pmk_set_view_mask -> K0m, pmk_probe_mask, the gate in Optim::postProcess) through the
C ABI against the reference's own answers on the same scene (tests/golden/config1_half_mask.npz).  Integer work: bit-exact."""
import os

import numpy as np
import pytest

from test_mask_cpu import GOLD, gold, gold_mask, masked_scene  # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mctx(masked_scene):
    from mvskit_b200 import pmk
    c = pmk.Context(nviews=masked_scene.nviews)
    c.set_scene(masked_scene.P, masked_scene.images, masks=masked_scene.masks)
    yield c
    c.close()


def test_mask_pyramid_bit_exact(mctx, gold, masked_scene):
    nlevels = 4
    for v in range(masked_scene.nviews):
        for l in range(nlevels):
            w, h = mctx.level_dims(v, l)
            want, got = gold_mask(gold, v, l, w), mctx.level_mask(v, l)
            assert (want is None) == (got is None), (v, l)
            if want is not None:
                assert np.array_equal(want, got), (v, l)


def test_get_mask_bit_exact(mctx, gold, masked_scene):
    pts = gold["points"]
    assert np.array_equal(mctx.probe_mask(pts), gold["getmask_all"])
    for v in range(masked_scene.nviews):
        assert np.array_equal(mctx.probe_mask(pts, view=v), gold["getmask_view"][v]), v


def test_post_process_gate(mctx, gold, masked_scene):
    from mvskit_b200 import pmk
    c, n, scal, views, nviews = (gold[k] for k in ("post_coord", "post_normal", "post_scal", "post_views", "post_nviews"))
    ret, images, nimg, grids, tmp = mctx.post_process(c, n, scal[:, 0].copy(), views, nviews)
    hit = gold["post_getmask"] == 0
    assert (ret[hit] == -1).all() and (nimg[hit] == 0).all()                     # the gate itself: exact
    # candidates that pass the gate behave exactly as on a context without masks ...
    plain = pmk.Context(nviews=masked_scene.nviews)
    plain.set_scene(masked_scene.P, masked_scene.images)
    pret, pimages, pnimg, pgrids, ptmp = plain.post_process(c, n, scal[:, 0].copy(), views, nviews)
    plain.close()
    assert (pret[hit] == 0).mean() > 0.9                                          # ... where the gate is what rejects them
    assert np.array_equal(ret[~hit], pret[~hit]) and np.array_equal(images[~hit], pimages[~hit]) and np.array_equal(grids[~hit], pgrids[~hit])
    # ... and as the reference's postProcess on the masked scene (list decisions within the NCC tolerance, as in test_cand_gpu)
    near = 0
    for i in np.nonzero(~hit)[0]:
        same = ret[i] == gold["post_ret"][i] and (ret[i] != 0 or (nimg[i] == gold["post_nimages"][i] and
                                                                    np.array_equal(images[i, :nimg[i]], gold["post_images"][i, :nimg[i]])))
        if not same:
            near += 1
            continue
        if ret[i] == 0:
            assert np.array_equal(grids[i, :nimg[i]], gold["post_grids"][i, :nimg[i]]), i
            assert abs(tmp[i] - gold["post_tmp"][i]) <= 1e-5
    assert near <= max(1, int((~hit).sum()) // 50), near


def test_sweep_never_creates_a_patch_outside_the_masks(mctx, masked_scene):
    """Every patch Propagate::run adds went through postProcess, hence through the gate: seeded only inside the silhouettes, the store
    holds no patch with PhotoSet::getMask == 0 after a sweep, while the same sweep without masks spreads over the whole plane."""
    from mvskit_b200 import pmk, synth
    coord, normal, scal, images, nimages = synth.seed_arrays(masked_scene, stride=4)
    inside = mctx.probe_mask(coord) != 0
    assert 50 < inside.sum() < len(coord)
    args = (coord[inside], normal[inside], scal[inside], images[inside], nimages[inside])
    mctx.store_clear()
    mctx.store_add(*args)
    mctx.propagate(0, 1234)
    got = mctx.store_get()
    assert got.n > 2 * inside.sum()
    assert (mctx.probe_mask(got.coord[:got.n]) != 0).all()
    plain = pmk.Context(nviews=masked_scene.nviews)
    plain.set_scene(masked_scene.P, masked_scene.images)
    plain.store_clear()
    plain.store_add(*args)
    plain.propagate(0, 1234)
    allp = plain.store_get()
    plain.close()
    assert (mctx.probe_mask(allp.coord[:allp.n]) == 0).sum() > 0.2 * allp.n
    mctx.store_clear()
