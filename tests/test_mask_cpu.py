"""Silhouette masks on the CPU side: the C restatement (oracle/pm_oracle.c: pmo_set_mask / pmo_get_mask*) and the host mirror's
PGM / PBM reader against what the reference itself answered on the same scene (tests/golden/config1_half_mask.npz, written by
tests/golden/make_golden_mask.py from oracle/_ref/libpmref.so).  Integer work: everything is compared exactly."""
import os
import subprocess
import sys

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1_half_mask.npz")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def masked_scene(gold):
    from mvskit_b200 import synth
    from make_golden_mask import scene_hash
    sc = synth.make_scene(1, scale=0.5).render().make_masks()
    assert scene_hash(sc) == str(gold["scene_sha256"]), "the scene generator drifted: regenerate tests/golden/config1_half_mask.npz"
    return sc


def gold_mask(gold, v, l, w):
    key = f"mask_v{v}_l{l}"
    if key not in gold.files:
        return None
    return (np.unpackbits(gold[key], axis=1)[:, :w] * 255).astype(np.uint8)


def test_oracle_mask_pyramid_and_lookups_equal_the_reference(gold, masked_scene):
    from oracle import pyoracle
    pyoracle.build(ref=False)
    co = pyoracle.COracle(masked_scene.P, masked_scene.images, masks=masked_scene.masks)
    assert list(gold["has_mask"]) == [1, 1, 0, 1, 1]
    for v in range(masked_scene.nviews):
        for l in range(co.nlevels):
            w, h = co.image_dims(v, l)
            want, got = gold_mask(gold, v, l, w), co.mask_level(v, l)
            assert (want is None) == (got is None), (v, l)
            if want is not None:
                assert want.shape == (h, w) and np.array_equal(want, got), (v, l)
    pts = gold["points"]
    assert np.array_equal(co.get_mask(pts), gold["getmask_all"])                       # PhotoSet::getMask(coord, m_level)
    for v in range(masked_scene.nviews):
        assert np.array_equal(co.get_mask(pts, view=v), gold["getmask_view"][v]), v    # Photo::getMask
    assert set(np.unique(gold["getmask_view"])) == {-1, 0, 255} and (gold["getmask_view"][2] == -1).all()
    assert np.array_equal(co.get_mask(gold["post_coord"]), gold["post_getmask"])


def test_reference_post_process_rejects_every_masked_out_candidate(gold):
    """optim.cpp:265: the gate sits before any NCC work, so getMask == 0 implies postProcess == -1 with m_images untouched by it."""
    hit = gold["post_getmask"] == 0
    assert hit.sum() > 100 and (~hit).sum() > 100
    assert (gold["post_ret"][hit] == -1).all()
    assert (gold["post_ret"][~hit] == 0).mean() > 0.9


def test_host_mirror_reads_pgm_and_pbm_like_the_reference(gold, masked_scene, tmp_path):
    """PhotoSet::readMask (mvskit_b200/host) on the files write_scene stores, thresholded as Image::alloc does (image.cpp:149-156),
    equals the reference's level-0 mask -- including the P4 file, whose bits the reference reads as one unpadded stream."""
    from mvskit_b200 import build, synth
    build.build()
    exe = build.build_host()
    prefix = synth.write_scene(masked_scene, str(tmp_path / "scene"), with_seeds=False)
    for v in range(masked_scene.nviews):
        out = str(tmp_path / f"m{v}.pgm")
        r = subprocess.run([exe, "--mask-io", prefix + "mask/%08d" % v, out], capture_output=True, text=True)
        if not gold["has_mask"][v]:
            assert r.returncode == 1
            continue
        assert r.returncode == 0, r.stderr
        raw = open(out, "rb").read()
        head = b"P5\n%d %d\n255\n" % (masked_scene.width, masked_scene.height)
        assert raw.startswith(head)
        grey = np.frombuffer(raw[len(head):], np.uint8).reshape(masked_scene.height, masked_scene.width)
        assert np.array_equal(np.where(grey > 127, 255, 0), gold_mask(gold, v, 0, masked_scene.width)), v
