"""JPEG ingest: pmk_set_view_jpeg decodes with nvJPEG on the device and feeds the same K0 pyramid.
JPEG decoders are not bit-reproducible against each other (the reference's own path is CImg + ImageMagick, third party and
absent), so the check is against libjpeg (PIL) on 4:4:4 streams: mean pixel difference below 0.6 grey levels (isolated pixels up to 6: different IDCT / colour-conversion rounding), INCC scores within 2e-2 (median 1e-3)."""
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_jpeg_views_match_libjpeg_and_score_like_the_raw_pixels(small_scene):
    PIL = pytest.importorskip("PIL.Image")
    from mvskit_b200 import pmk
    scene = small_scene
    jpegs, decoded = [], []
    for im in scene.images:
        buf = io.BytesIO()
        PIL.fromarray(im).save(buf, format="JPEG", quality=97, subsampling=0)
        jpegs.append(buf.getvalue())
        decoded.append(np.asarray(PIL.open(io.BytesIO(buf.getvalue())).convert("RGB")))
    cj = pmk.Context(nviews=scene.nviews)
    for v in range(scene.nviews):
        w, h = cj.set_view_jpeg(v, scene.P[v], jpegs[v])
        assert (w, h) == (scene.images[v].shape[1], scene.images[v].shape[0])
    cp = pmk.Context(nviews=scene.nviews)
    cp.set_scene(scene.P, decoded)                       # the same streams decoded by libjpeg
    for v in (0, scene.nviews - 1):
        a, b = cj.level_image(v, 0).astype(int), decoded[v].astype(int)
        assert np.abs(a - b).max() <= 6 and np.abs(a - b).mean() < 0.6, (np.abs(a - b).max(), np.abs(a - b).mean())
    c, n, vw, nv = scene.hypotheses(2048, seed=77, well_observed=True)
    ij, _ = cj.ncc_eval(c, n, vw, nv)
    ip, _ = cp.ncc_eval(c, n, vw, nv)
    assert np.array_equal(ij == 2.0, ip == 2.0)
    ok = ip != 2.0
    assert np.abs(ij[ok] - ip[ok]).max() <= 2e-2 and np.median(np.abs(ij[ok] - ip[ok])) <= 1e-3
    with pytest.raises(pmk.PmkError, match="not a JPEG"):
        pmk.Context(nviews=1).set_view_jpeg(0, scene.P[0], b"P6 not a jpeg at all")
    cj.close(); cp.close()
