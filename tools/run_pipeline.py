#!/usr/bin/env python
"""Run init -> (propagate, filter, updateThreshold) x ITER on one GPU for a synthetic config and print per-stage numbers.
   python tools/run_pipeline.py --config 1 --scale 0.5 --iters 3 --group 5"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvskit_b200 import pmk, synth  # noqa: E402


def _render_worker(args):
    config, scale, nviews, views = args
    sc = synth.make_scene(config, scale=scale, nviews=nviews)
    sc.render(views=views)
    return [(v, sc.images[v]) for v in views]


def render_parallel(scene, config, scale, nviews, procs):
    """scene.render() spread over `procs` processes (every view renders independently and deterministically)."""
    import multiprocessing as mp
    V = scene.nviews
    jobs = [(config, scale, nviews, list(range(i, V, procs))) for i in range(procs)]
    scene.images = [None] * V
    with mp.get_context("spawn").Pool(procs) as pool:
        for part in pool.map(_render_worker, jobs):
            for v, im in part:
                scene.images[v] = im
    return scene


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--scale", type=float, default=0.5)
    ap.add_argument("--nviews", type=int, default=None)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--group", type=int, default=1)
    ap.add_argument("--cell-capacity", type=int, default=0)
    ap.add_argument("--procs", type=int, default=1, help="processes rendering the synthetic views")
    ap.add_argument("--seed-stride", type=int, default=4, help="one seed every this many cells")
    a = ap.parse_args()
    t0 = time.time()
    scene = synth.make_scene(a.config, scale=a.scale, nviews=a.nviews)
    scene = render_parallel(scene, a.config, a.scale, a.nviews, a.procs) if a.procs > 1 else scene.render()
    print(f"rendered {scene.nviews} views {scene.width}x{scene.height} in {time.time() - t0:.1f} s", flush=True)
    ctx = pmk.Context(nviews=scene.nviews, sweep_group=a.group, cell_capacity=a.cell_capacity)
    ctx.set_scene(scene.P, scene.images)
    t0 = time.time()
    coord, normal, scal, images, nimg = synth.seed_arrays(scene, stride=a.seed_stride)
    print(f"{len(coord)} seeds in {time.time() - t0:.1f} s", flush=True)
    ctx.set_depth(0); ctx.store_clear(); ctx.store_add(coord, normal, scal, images, nimg); ctx.set_depth(1)
    print("seeds", len(coord), "views", scene.nviews, "grid", ctx.grid_dims(0), "group", a.group, flush=True)
    T0 = time.time()
    for it in range(a.iters):
        t = time.time(); st = ctx.propagate(it, 0x5EED0001); ctx.sync(); t1 = time.time() - t
        nprop = ctx.store_count()
        ck_prop = ctx.store_checksum()
        t = time.time(); c = ctx.filter(); ctx.sync(); t2 = time.time() - t
        print(f"iter {it}: checksum after propagate {ck_prop}, after filter {ctx.store_checksum()}", flush=True)
        ctx.update_threshold()
        g = ctx.store_get()
        q = None
        if a.config == 1:
            q = np.quantile(np.abs(g.coord[:, 2]) / scene.scene_scale, [0.5, 0.9, 0.99])
        elif a.config == 2:
            r = np.linalg.norm(g.coord[:, :3], axis=1)
            on_sphere = np.abs(r - 1.0) < 0.05
            q = (float(on_sphere.mean()), np.quantile(np.abs(r[on_sphere] - 1.0) / scene.scene_scale, [0.5, 0.9]))
        print(f"iter {it}: propagate {t1:.2f}s -> {nprop} patches, {st['evals'] / t1 / 1e6:.1f} M evals/s, slowest-cell sum {st["step_max_ns"] / 1e9:.2f}s of {st["steps"]} steps, cell time {st["cell_ns"] / 1e9:.1f}s, stats {st}; filter {t2:.3f}s {c}; "
              f"quality {q}; launches {ctx.launch_count()}", flush=True)
    dt = time.time() - T0
    print(f"total {dt:.2f} s; {g.n} patches; {g.n / dt:.0f} patches/s")


if __name__ == "__main__":
    main()
