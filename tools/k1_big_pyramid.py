#!/usr/bin/env python
"""K1 on a pyramid that does not fit L2 (config 3: 49 views 1600x1200, 1.0 GB; config 5: 128 views 1920x1080, 2.8 GB): device-resident
launches with the L2 flushed before each, for timing and for one `ncu --set full -k regex:k1_ncc --launch-skip 3 --launch-count 1` capture.
   python tools/k1_big_pyramid.py [--config 3] [--log2n 19] [--steps 5]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mvskit_b200 import pmk  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--log2n", type=int, nargs="+", default=[19])
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    scene = bench.get_scene(a.config, 1.0)
    ctx = pmk.Context(nviews=scene.nviews)
    ctx.set_scene(scene.P, scene.images)
    for lg in a.log2n:
        c, n, v, nv = bench.get_hypotheses(scene, 1 << lg, 7, a.config, 1.0, "grid", procs=bench.host_procs(1))
        bufs = [ctx.alloc(x.nbytes).upload(x) for x in (c, n, v, nv)]
        oi, on = ctx.alloc(len(c) * 4), ctx.alloc(len(c) * 4)
        for _ in range(3):
            ctx.ncc_eval_dev(len(c), bufs[0], bufs[1], bufs[2], bufs[3], v.shape[1], oi, on)
        ms = []
        for _ in range(a.steps):
            ctx.flush_l2()
            ctx.timer_begin()
            ctx.ncc_eval_dev(len(c), bufs[0], bufs[1], bufs[2], bufs[3], v.shape[1], oi, on)
            ms.append(ctx.timer_end())
        print(f"config {a.config}: 2^{lg} hypotheses, {np.mean(ms):.3f} ms per launch, {len(c) / np.mean(ms) / 1e3:.1f} M evals/s "
              f"(valid views per eval {float(np.minimum(nv, 6).mean()):.2f})", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
