"""Golden vectors generated from the reference itself (tests/golden/make_golden.py).  CPU part: the plain-C
restatement reproduces them bit for bit.  GPU part: the product meets the north-star bar against them."""
import hashlib
import os

import numpy as np
import pytest

from conftest import assert_bits_equal

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1_half.npz"))


def test_scene_generator_has_not_drifted(small_scene):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(small_scene.P).tobytes())
    for im in small_scene.images:
        h.update(np.ascontiguousarray(im).tobytes())
    assert h.hexdigest() == str(G["scene_sha256"]), "synthetic scene changed: regenerate tests/golden with make_golden.py"


def test_c_oracle_reproduces_golden(coracle):
    c, n, vw, nv = G["coord"], G["normal"], G["views"], G["nviews"]
    v0 = vw[:, 0].copy()
    assert_bits_equal(coracle.project(v0, c), G["project"], "project")
    assert_bits_equal(coracle.get_unit(v0, c), G["unit"], "unit")
    px, py = coracle.get_paxes(v0, c, n)
    assert_bits_equal(px, G["px"], "px"); assert_bits_equal(py, G["py"], "py")
    ixy, ok = coracle.cells(v0, c)
    assert np.array_equal(ixy, G["cell"]) and np.array_equal(ok, G["cell_ok"])
    incc, ncc, lv = coracle.compute_ncc(c, n, vw, nv, True)
    assert_bits_equal(incc, G["incc"], "incc"); assert_bits_equal(ncc, G["ncc"], "ncc")
    assert np.array_equal(lv, G["levels"])
    for i in range(0, len(c), 5):
        assert_bits_equal(coracle.set_inccs(c[i], n[i], G["all_views"][i], 0), G["inccs_1vsall"][i], "inccs")
        assert_bits_equal(coracle.set_inccs_pair(c[i], n[i], G["all_views"][i], 1), G["inccs_pair_robust"][i], "inccs pair")


@pytest.mark.gpu
def test_product_meets_bar_on_golden(small_scene):
    from mvskit_b200 import pmk
    ctx = pmk.Context(nviews=small_scene.nviews)
    ctx.set_scene(small_scene.P, small_scene.images)
    c, n, vw, nv = G["coord"], G["normal"], G["views"], G["nviews"]
    v0 = vw[:, 0].copy()
    pr = ctx.probe(v0, c, n)
    assert_bits_equal(pr["project"], G["project"], "project")
    assert_bits_equal(pr["unit"], G["unit"], "unit")
    assert_bits_equal(pr["px"], G["px"], "px"); assert_bits_equal(pr["py"], G["py"], "py")
    assert np.array_equal(pr["cell"], G["cell"]) and np.array_equal(pr["cell_ok"], G["cell_ok"])
    incc, ncc, lv = ctx.ncc_eval(c, n, vw, nv, want_levels=True)
    assert np.array_equal(lv, G["levels"])
    assert np.array_equal(incc == 2.0, G["incc"] == 2.0)
    ok = G["incc"] != 2.0
    assert np.abs(incc[ok] - G["incc"][ok]).max() <= 1e-4
    allv = G["all_views"]
    nall = np.full(len(c), allv.shape[1], np.int32)
    one = ctx.set_inccs(c, n, allv, nall, 0)
    assert np.array_equal(one == 2.0, G["inccs_1vsall"] == 2.0)
    ok = G["inccs_1vsall"] != 2.0
    assert np.abs(one[ok] - G["inccs_1vsall"][ok]).max() <= 1e-4
    pair = ctx.set_inccs(c, n, allv, nall, 1, pairwise=True)
    ok = G["inccs_pair_robust"] != 2.0
    assert np.array_equal(pair == 2.0, ~ok)
    assert np.abs(pair[ok] - G["inccs_pair_robust"][ok]).max() <= 1e-4
    ret, images, nimg, ds, asc = ctx.pre_process(c, n, vw, nv)
    same = (ret == G["pre_ret"]) & (nimg == G["pre_nimages"])
    assert same.mean() >= 0.99
    for i in np.nonzero(same)[0]:
        assert np.array_equal(images[i, :nimg[i]], G["pre_images"][i, :nimg[i]])
        if nimg[i] > 0:
            assert_bits_equal(ds[i:i + 1], G["pre_scal"][i:i + 1, 1], "dscale")
    keep = G["refine_keep"]
    gc, gn, gncc, gtr = ctx.refine(c[keep], n[keep], G["pre_scal"][keep, 1], G["pre_images"][keep], G["pre_nimages"][keep],
                                   G["refine_streams"], int(G["refine_seed"]), trace=True)
    err = np.linalg.norm(gc[:, :3] - G["refine_coord"][:, :3], axis=1) / small_scene.scene_scale
    assert np.quantile(err, 0.95) <= 1e-3
    assert np.abs(gtr[:, 0, :3] - G["refine_trace"][:, 0, :3]).max() <= 1e-5
    ctx.close()
