#!/usr/bin/env python
"""Generate tests/golden/config1_half_mask.npz from the reference ITSELF (oracle/_ref/libpmref.so) on config 1 at half size WITH
silhouette masks (mvskit_b200.synth.Scene.make_masks: grey PGM on views 0, 3, 4, binary PBM on view 1, no mask on view 2):

    python tests/golden/make_golden_mask.py

Holds what the reference's own code answers on that scene: Image::m_masks of every view and level (after Image::alloc's
threshold and buildMaskPyramid), PhotoSet::getMask(coord, m_level) and PhotoSet::getMask(view, coord, m_level) on a batch of
points, and Optim::postProcess (whose second statement is the mask gate, optim.cpp:265) on post-refine candidates.
The reference keeps one scene per process, so this runs in a process of its own and the tests read the file.
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mvskit_b200 import synth          # noqa: E402
from oracle import pyoracle            # noqa: E402


def scene_hash(scene) -> str:
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(scene.P).tobytes())
    for im in scene.images:
        h.update(np.ascontiguousarray(im).tobytes())
    for m in scene.masks:
        h.update(b"-" if m is None else np.ascontiguousarray(m).tobytes())
    return h.hexdigest()


def mask_points(scene, n=4096, seed=31):
    """surface points (many land on mask rims), plus points far off the surface / behind the cameras"""
    c = scene.hypotheses(n, seed=seed, well_observed=False)[0]
    rng = np.random.default_rng(seed + 1)
    far = np.ones((n // 8, 4), np.float32)
    far[:, :3] = rng.uniform(-6.0, 6.0, size=(n // 8, 3))
    return np.concatenate([c, far]).astype(np.float32)


def main():
    pyoracle.build(ref=True)
    scene = synth.make_scene(1, scale=0.5).render().make_masks()
    prefix = synth.write_scene(scene, tempfile.mkdtemp(prefix="pm_golden_mask_"))
    ref = pyoracle.RefLib(prefix)
    out = dict(scene_sha256=np.array(scene_hash(scene)))
    has = np.zeros(scene.nviews, np.int32)
    for v in range(scene.nviews):
        for l in range(ref.nlevels):
            m = ref.mask_level(v, l)
            if m is not None:
                has[v] = 1
                out[f"mask_v{v}_l{l}"] = np.packbits(m > 0, axis=1)
                assert set(np.unique(m)) <= {0, 255}
    out["has_mask"] = has
    pts = mask_points(scene)
    out["points"] = pts
    out["getmask_all"] = ref.get_mask(pts)
    out["getmask_view"] = np.stack([ref.get_mask(pts, view=v) for v in range(scene.nviews)])
    # Optim::postProcess on candidates in their post-refine state (the recipe of tests/test_cand_gpu.py::test_post_process)
    c, n, vw, nv = scene.hypotheses(768, seed=21, depth_jitter=0.004, normal_jitter_deg=12.0)
    rret, rb = ref.pre_process(c, n, vw, nv)
    keep = np.nonzero(rret == 0)[0]
    c, n = c[keep], n[keep]
    views, nviews = rb.images[keep], rb.nimages[keep]
    incc, ncc = ref.compute_ncc(c, n, views, nviews)
    scal = np.zeros((len(keep), 4), np.float32)
    scal[:, 0], scal[:, 1], scal[:, 2] = ncc, rb.scal[keep, 1], rb.scal[keep, 2]
    ref.set_depth(0)
    pret, pb = ref.post_process(c, n, scal, views, nviews)
    out.update(post_coord=c, post_normal=n, post_scal=scal, post_views=views, post_nviews=nviews, post_ret=pret,
               post_images=pb.images, post_nimages=pb.nimages, post_grids=pb.grids, post_tmp=pb.scal[:, 3],
               post_getmask=ref.get_mask(c))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config1_half_mask.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; scene", out["scene_sha256"])
    print("getMask(all):", dict(zip(*np.unique(out["getmask_all"], return_counts=True))), " postProcess ret:",
          dict(zip(*np.unique(pret, return_counts=True))), " gate hits:", int((out["post_getmask"] == 0).sum()), "of", len(keep))


if __name__ == "__main__":
    main()
