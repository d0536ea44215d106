#!/usr/bin/env python
"""Smoke the less-travelled configurations of the pipeline on a small scene: window sizes 5 / 9 / 11, Philox jitter, csize 1,
level 0/2, minImageNum 2/4: one propagate + filter each, sanity of the result."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvskit_b200 import pmk, synth  # noqa: E402


def run(scene, **kw):
    ctx = pmk.Context(nviews=scene.nviews, sweep_group=scene.nviews, **kw)
    ctx.set_scene(scene.P, scene.images)
    seeds = synth.seed_arrays(scene)
    ctx.set_depth(0); ctx.store_clear(); ctx.store_add(*seeds); ctx.set_depth(1)
    st = ctx.propagate(0, 7)
    c = ctx.filter()
    ctx.update_threshold()
    st2 = ctx.propagate(1, 7)
    c2 = ctx.filter()
    g = ctx.store_get()
    z = np.quantile(np.abs(g.coord[:, 2]) / scene.scene_scale, [0.5, 0.9]) if g.n else None
    ctx.close()
    return len(seeds[0]), c, c2, z


def main():
    scene = synth.make_scene(1, scale=0.5).render()
    for kw in (dict(wsize=5), dict(wsize=9), dict(wsize=11), dict(jitter_mode=1), dict(min_image_num=2), dict(min_image_num=4), dict(csize=1)):
        try:
            print(kw, run(scene, **kw), flush=True)
        except Exception as exc:
            print(kw, "ERROR", exc, flush=True)


if __name__ == "__main__":
    main()
