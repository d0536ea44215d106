#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference ITSELF (oracle/_ref/libpmref.so = the unmodified sources under
/root/reference compiled against oracle/shim/).  Run where /root/reference exists:

    python tests/golden/make_golden.py

The scene is regenerated from its seed (mvskit_b200.synth, config 1 at half size); a SHA-256 of the rendered images
is stored so a drift of the generator is detected instead of silently invalidating the vectors.
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
    return h.hexdigest()


def main():
    pyoracle.build(ref=True)
    scene = synth.make_scene(1, scale=0.5).render()
    prefix = synth.write_scene(scene, tempfile.mkdtemp(prefix="pm_golden_"))
    ref = pyoracle.RefLib(prefix)
    a = scene.hypotheses(192, seed=101, well_observed=True)
    b = scene.hypotheses(64, seed=102, well_observed=False, normal_jitter_deg=35.0)
    c, n, vw, nv = (np.concatenate([x, y]) for x, y in zip(a, b))
    v0 = vw[:, 0].copy()
    out = dict(scene_sha256=np.array(scene_hash(scene)), coord=c, normal=n, views=vw, nviews=nv)
    out["project"] = ref.project(v0, c)
    out["unit"] = ref.get_unit(v0, c)
    out["px"], out["py"] = ref.get_paxes(v0, c, n)
    out["cell"], out["cell_ok"] = ref.cells(v0, c)
    out["incc"], out["ncc"] = ref.compute_ncc(c, n, vw, nv)
    lv = np.full((len(c), ref.tau), -1, np.int32)
    for i in range(len(c)):
        for k in range(min(nv[i], ref.tau)):
            lv[i, k] = ref.get_tex(c[i], n[i], int(vw[i, 0]), int(vw[i, k]))[2]
    out["levels"] = lv
    V = scene.nviews
    allv = np.array([[vw[i, 0]] + [v for v in range(V) if v != vw[i, 0]] for i in range(len(c))], np.int32)
    out["all_views"] = allv
    out["inccs_1vsall"] = np.stack([ref.set_inccs(c[i], n[i], allv[i], 0) for i in range(len(c))])
    out["inccs_pair_robust"] = np.stack([ref.set_inccs_pair(c[i], n[i], allv[i], 1) for i in range(len(c))])
    pret, pb = ref.pre_process(c, n, vw, nv)
    out["pre_ret"], out["pre_images"], out["pre_nimages"], out["pre_scal"] = pret, pb.images, pb.nimages, pb.scal
    keep = np.nonzero(pret == 0)[0][:48]
    streams = np.arange(len(keep), dtype=np.uint64) * np.uint64(31) + np.uint64(5)
    seed = 0xC0FFEE1234567890
    rc, rn, rncc, rtr = ref.refine(c[keep], n[keep], pb.scal[keep, 1], pb.images[keep], pb.nimages[keep], streams, seed, trace=True)
    out.update(refine_keep=keep, refine_streams=streams, refine_seed=np.array(seed, np.uint64), refine_coord=rc, refine_normal=rn,
               refine_ncc=rncc, refine_trace=rtr)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config1_half.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; scene", out["scene_sha256"])


if __name__ == "__main__":
    main()
