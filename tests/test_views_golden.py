"""Parity at the FULL VIEW COUNTS of BASELINE configs 2, 3 and 5 (47 / 49 / 128 views, reduced image size) against golden vectors
generated from the reference itself (tests/golden/make_golden_views.py): the O(nimages) paths -- Optim::setINCCs over every view,
preProcess / postProcess view lists, addPatch registrations, depth maps and visible lists of every view -- and the limits that sit on
them (nviews <= 128, the candidate kernels' view lists, cell_capacity = max(96, 4 nviews)).  CPU part: the C restatement reproduces
the scores bit for bit.  GPU part: the product, through the C ABI."""
import hashlib
import os

import numpy as np
import pytest

from conftest import assert_bits_equal

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {2: 0.25, 3: 0.1, 5: 0.125}
_SCENES = {}


def _golden(config):
    return np.load(os.path.join(HERE, "golden", f"config{config}_views.npz"))


def _scene(config):
    if config not in _SCENES:
        from mvskit_b200 import synth
        sc = synth.make_scene(config, scale=CASES[config]).render()
        h = hashlib.sha256()
        h.update(np.ascontiguousarray(sc.P).tobytes())
        for im in sc.images:
            h.update(np.ascontiguousarray(im).tobytes())
        assert h.hexdigest() == str(_golden(config)["scene_sha256"]), "synthetic scene changed: regenerate with make_golden_views.py"
        _SCENES[config] = sc
    return _SCENES[config]


@pytest.mark.parametrize("config", [2, 3])
def test_c_oracle_reproduces_the_reference_at_full_view_count(config):
    from oracle import pyoracle
    pyoracle.build(ref=False)
    G, sc = _golden(config), _scene(config)
    orc = pyoracle.COracle(sc.P, sc.images)
    c, n, vw, nv = G["coord"], G["normal"], G["views"], G["nviews"]
    incc, ncc, lv = orc.compute_ncc(c, n, vw, nv, True)
    assert_bits_equal(incc, G["incc"], "incc"); assert_bits_equal(ncc, G["ncc"], "ncc")
    assert np.array_equal(lv, G["levels"])
    for i in range(0, len(c), 4):
        assert_bits_equal(orc.set_inccs(c[i], n[i], G["all_views"][i], 0), G["inccs_1vsall"][i], f"setINCCs over {sc.nviews} views")


@pytest.mark.gpu
@pytest.mark.parametrize("config", [2, 3, 5])
def test_product_at_full_view_count(config):
    from mvskit_b200 import pmk
    G, sc = _golden(config), _scene(config)
    V = sc.nviews
    ctx = pmk.Context(nviews=V)
    ctx.set_scene(sc.P, sc.images)
    c, n, vw, nv = G["coord"], G["normal"], G["views"], G["nviews"]
    # ---- K1: levels / validity exact, scores within 1e-4 ----
    incc, ncc, lv = ctx.ncc_eval(c, n, vw, nv, want_levels=True)
    assert np.array_equal(lv, G["levels"])
    assert np.array_equal(incc == 2.0, G["incc"] == 2.0)
    ok = G["incc"] != 2.0
    assert np.abs(incc[ok] - G["incc"][ok]).max() <= 1e-4
    # ---- setINCCs 1-vs-all over ALL views ----
    allv = G["all_views"]
    one = ctx.set_inccs(c, n, allv, np.full(len(c), V, np.int32), 0)
    assert np.array_equal(one == 2.0, G["inccs_1vsall"] == 2.0), f"validity of {V} views per patch"
    ok = G["inccs_1vsall"] != 2.0
    assert np.abs(one[ok] - G["inccs_1vsall"][ok]).max() <= 1e-4
    # ---- preProcess from the bare reference view: the ordered view list is integer work ----
    ret, images, nimg, ds, asc = ctx.pre_process(c, n, vw[:, :1].copy(), np.ones(len(c), np.int32))
    same = (ret == G["pre_ret"]) & (nimg == G["pre_nimages"])
    lim = 1.0 - 0.4                                                # constraintImages at m_nccThresholdBefore (optim.cpp:207-219)
    for i in np.nonzero(~same)[0]:                                 # a list may only differ where a view's INCC sits on the threshold
        assert np.min(np.abs(G["inccs_1vsall"][i][1:] - lim)) <= 1e-4, (i, ret[i], G["pre_ret"][i])
    assert same.mean() >= 0.97, same.mean()
    for i in np.nonzero(same)[0]:
        assert np.array_equal(images[i, :nimg[i]], G["pre_images"][i, :nimg[i]]), i
        if nimg[i] > 0:
            assert_bits_equal(ds[i:i + 1], G["pre_scal"][i:i + 1, 1], "m_dscale")
    # ---- postProcess on the reference's preProcess result (store-free part, m_depth 0) ----
    idx = G["post_index"]
    ctx.set_depth(0)
    pret, pimg, pn, pgrids, ptmp = ctx.post_process(c[idx], n[idx], G["ncc"][idx], G["pre_images"][idx], G["pre_nimages"][idx])
    psame = (pret == G["post_ret"]) & ((pn == G["post_nimages"]) | (pret != 0))
    assert psame.mean() >= 0.97, psame.mean()
    lists_equal = 0
    for i in np.nonzero(psame & (pret == 0))[0]:
        k = pn[i]
        if np.array_equal(pimg[i, :k], G["post_images"][i, :k]):
            lists_equal += 1
            assert np.array_equal(pgrids[i, :k], G["post_grids"][i, :k]), i          # setGrids, bit-exact cells
            assert abs(ptmp[i] - G["post_tmp"][i]) <= 1e-4 * max(1.0, abs(G["post_tmp"][i]))
    assert lists_equal >= 0.97 * int((psame & (pret == 0)).sum()), (lists_equal, int((psame & (pret == 0)).sum()))   # setRefImage: argmin of INCC sums
    # ---- the store at this view count: registrations in every view, depth maps, visible lists ----
    ctx.store_clear()
    ctx.store_add(G["seed_coord"], G["seed_normal"], G["seed_scal"], G["seed_images"], G["seed_nimages"])
    g = ctx.store_get()
    assert g.n == len(G["seed_coord"])
    assert_bits_equal(g.coord, G["seed_coord"], "collect order")
    assert np.array_equal(g.nimages, G["seed_nimages"])
    for i in range(0, g.n, 7):
        k = g.nimages[i]
        assert np.array_equal(g.grids[i, :k], G["seed_grids"][i, :k]), i
    ctx.set_depth(1)
    assert ctx.filter_rebuild(0) == g.n
    gb = ctx.store_get()
    assert_bits_equal(gb.coord, G["rebuilt_coord"], "collect order after the rebuild")
    assert np.array_equal(gb.nvimages, G["rebuilt_nvimages"])
    for i in range(0, gb.n, 5):
        k = gb.nvimages[i]
        assert np.array_equal(gb.vimages[i, :k], G["rebuilt_vimages"][i, :k]) and np.array_equal(gb.vgrids[i, :k], G["rebuilt_vgrids"][i, :k]), i
    for v in range(V):
        assert np.array_equal(ctx.store_depth_map(v), G["depth_maps"][v]), f"m_dpgrids of view {v} of {V}"
        assert np.array_equal(ctx.store_cell_counts(v, 0), G["pcounts"][v]), f"m_pgrids sizes of view {v}"
        assert np.array_equal(ctx.store_cell_counts(v, 1), G["vcounts"][v]), f"m_vpgrids sizes of view {v}"
    ctx.close()
