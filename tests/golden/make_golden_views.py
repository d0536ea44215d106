#!/usr/bin/env python
"""Golden vectors at the FULL VIEW COUNTS of BASELINE configs 2, 3 and 5 (47 / 49 / 128 views) at reduced image size, generated
from the reference ITSELF (oracle/_ref/libpmref.so).  This is where the O(nimages) paths bite: Optim::setINCCs 1-vs-all over every
view (optim.cpp:708-746), preProcess / postProcess view lists (:137-163, :260-298), PatchManager::addPatch registrations and the
depth maps / visible lists of every view (patch_manager.cpp:158-221, 267-301), view-list and cell capacities.

    python tests/golden/make_golden_views.py          # one subprocess per config (the reference holds one scene per process)

The scenes are regenerated from their seeds at test time; a SHA-256 of the rendered images guards against generator drift."""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CASES = {2: 0.25, 3: 0.1, 5: 0.125}          # config -> image scale: 160x120 x 47, 160x120 x 49, 240x132 x 128


def scene_hash(scene) -> str:
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(scene.P).tobytes())
    for im in scene.images:
        h.update(np.ascontiguousarray(im).tobytes())
    return h.hexdigest()


def make(config: int):
    from mvskit_b200 import synth
    from oracle import pyoracle
    scale = CASES[config]
    scene = synth.make_scene(config, scale=scale).render()
    prefix = synth.write_scene(scene, tempfile.mkdtemp(prefix=f"pm_golden_v{config}_"), seed_stride=3)
    ref = pyoracle.RefLib(prefix)
    V = scene.nviews
    a = scene.hypotheses(96, seed=201, well_observed=True)
    b = scene.hypotheses(32, seed=202, well_observed=False, normal_jitter_deg=35.0)
    c, n, vw, nv = (np.concatenate([x, y]) for x, y in zip(a, b))
    out = dict(scene_sha256=np.array(scene_hash(scene)), scale=np.float32(scale), coord=c, normal=n, views=vw, nviews=nv)
    # K1: scores + pyramid level / validity of each sampled view
    out["incc"], out["ncc"] = ref.compute_ncc(c, n, vw, nv)
    lv = np.full((len(c), ref.tau), -1, np.int32)
    for i in range(len(c)):
        if nv[i] < 2:                        # computeINCC returns 2.0 before any getTex (optim.cpp:631,643): no level is ever chosen
            continue
        for k in range(min(nv[i], ref.tau)):
            lv[i, k] = ref.get_tex(c[i], n[i], int(vw[i, 0]), int(vw[i, k]))[2]
    out["levels"] = lv
    # setINCCs 1-vs-all over ALL views
    allv = np.array([[vw[i, 0]] + [v for v in range(V) if v != vw[i, 0]] for i in range(len(c))], np.int32)
    out["all_views"] = allv
    out["inccs_1vsall"] = np.stack([ref.set_inccs(c[i], n[i], allv[i], 0) for i in range(len(c))])
    # preProcess from the bare reference view, postProcess on its result
    ref.set_ncc_thresholds(0.7, 0.4)
    one = vw[:, :1].copy()
    pret, pb = ref.pre_process(c, n, one, np.ones(len(c), np.int32))
    out["pre_ret"], out["pre_images"], out["pre_nimages"], out["pre_scal"] = pret, pb.images, pb.nimages, pb.scal
    ok = np.nonzero(pret == 0)[0]
    scal = pb.scal[ok].copy()
    scal[:, 0] = out["ncc"][ok]
    ref.set_depth(0)
    qret, qb = ref.post_process(c[ok], n[ok], scal, pb.images[ok], pb.nimages[ok])
    out["post_index"], out["post_ret"], out["post_images"], out["post_nimages"], out["post_grids"], out["post_tmp"] = ok, qret, qb.images, qb.nimages, qb.grids, qb.scal[:, 3]
    # the store: seeds -> addPatch in every view, then Filter::setDepthMapsVGridsVPGridsAddPatchV(0) at m_depth 1
    ref.clear_patches(); ref.set_depth(0); ref.create_patches()
    seeds = ref.get_patches()
    out["seed_coord"], out["seed_normal"], out["seed_scal"], out["seed_images"], out["seed_nimages"] = seeds.coord, seeds.normal, seeds.scal, seeds.images, seeds.nimages
    out["seed_grids"] = seeds.grids
    ref.set_depth(1)
    ref.filter_rebuild(0)
    rb = ref.get_patches()
    out["rebuilt_coord"], out["rebuilt_nvimages"], out["rebuilt_vimages"], out["rebuilt_vgrids"] = rb.coord, rb.nvimages, rb.vimages, rb.vgrids
    out["depth_maps"] = np.stack([ref.depth_map(v) for v in range(V)])
    out["pcounts"] = np.stack([ref.cell_counts(v, 0) for v in range(V)])
    out["vcounts"] = np.stack([ref.cell_counts(v, 1) for v in range(V)])
    path = os.path.join(ROOT, "tests", "golden", f"config{config}_views.npz")
    np.savez_compressed(path, **out)
    print(f"config {config}: {V} views {scene.width}x{scene.height}, {len(c)} hypotheses ({int((pret == 0).sum())} pass preProcess, "
          f"{int((qret == 0).sum())} pass postProcess), {seeds.n} seeds, wrote {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        make(int(sys.argv[1]))
    else:
        from oracle import pyoracle
        pyoracle.build(ref=True)
        for cfg in CASES:
            subprocess.run([sys.executable, os.path.abspath(__file__), str(cfg)], check=True)
