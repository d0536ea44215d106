#!/usr/bin/env python
"""Drive every candidate / store / filter kernel once on a 47-view scene so that one `ncu --set full -k regex:...` pass can
capture K2 (k2_set_inccs), view selection (k_pre_process / k_post_process), K3 (k3_refine), the sweep and its apply kernels,
K5 (rebuild), K6..K9 (filters).   python tools/profile_kernels.py [--scale 0.5] [--ncand 8192]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvskit_b200 import pmk, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--scale", type=float, default=0.5)
    ap.add_argument("--ncand", type=int, default=8192)
    ap.add_argument("--iters", type=int, default=2)
    a = ap.parse_args()
    scene = synth.make_scene(a.config, scale=a.scale).render()
    V = scene.nviews
    ctx = pmk.Context(nviews=V, sweep_group=V)
    ctx.set_scene(scene.P, scene.images)
    c, n, vw, nv = scene.hypotheses(a.ncand, seed=11)
    # K2 over ALL views (the O(nimages) path of preProcess / postProcess): view list = reference + every other view
    allv = np.zeros((len(c), V), np.int32)
    for i in range(len(c)):
        rest = [v for v in range(V) if v != vw[i, 0]]
        allv[i] = [vw[i, 0]] + rest
    nall = np.full(len(c), V, np.int32)
    t = time.time(); ctx.set_inccs(c, n, allv, nall, 0); print(f"k2_set_inccs 1-vs-all over {V} views: {time.time() - t:.3f} s")
    t = time.time(); ctx.set_inccs(c[:1024], n[:1024], vw[:1024], nv[:1024], 1, pairwise=True); print(f"k2_set_inccs pairwise: {time.time() - t:.3f} s")
    t = time.time(); ret, imgs, nimg, dscale, ascale = ctx.pre_process(c, n, vw[:, :1].copy(), np.ones(len(c), np.int32)); print(f"k_pre_process: {time.time() - t:.3f} s, {int((ret == 0).sum())} pass")
    ok = np.nonzero(ret == 0)[0]
    streams = np.arange(len(ok), dtype=np.uint64)
    t = time.time(); rc, rn, rncc = ctx.refine(c[ok], n[ok], dscale[ok], imgs[ok], nimg[ok], streams, 0x5EED)[:3]; print(f"k3_refine {len(ok)} candidates: {time.time() - t:.3f} s")
    t = time.time(); ctx.post_process(rc, rn, rncc, imgs[ok], nimg[ok]); print(f"k_post_process: {time.time() - t:.3f} s")
    coord, normal, scal, images, nim = synth.seed_arrays(scene, stride=4)
    ctx.set_depth(0); ctx.store_clear(); ctx.store_add(coord, normal, scal, images, nim); ctx.set_depth(1)
    for it in range(a.iters):
        t = time.time(); st = ctx.propagate(it, 0x5EED0001); ctx.sync(); t1 = time.time() - t
        t = time.time(); cnt = ctx.filter(); ctx.sync(); t2 = time.time() - t
        ctx.update_threshold()
        print(f"iter {it}: propagate {t1:.2f} s, filter {t2:.3f} s {cnt}, checksum {ctx.store_checksum()}")
    print("launches", ctx.launch_count())


if __name__ == "__main__":
    main()
