#!/usr/bin/env python
"""How much of the sweep's wall time is the per-anti-diagonal barrier?  Records the time of every dest cell in one pass (config 2,
all views per step) and compares, per view and for the whole pass: the sum over steps of the slowest cell (what the barrier
costs now), the longest dependency chain when the barrier is only kept every M anti-diagonals, and the work bound (total / warps)."""
import os
import sys
import ctypes as C

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench  # noqa: E402
from mvskit_b200 import pmk, synth  # noqa: E402


def chain(T, M):
    """longest path through the DAG (x,y) <- (x-1,y), (x,y-1) when a barrier sits after every M anti-diagonals; per window the
    path restarts from zero.  T: [views, gh, gw] seconds.  Returns the sum over windows of the window's longest path."""
    V, gh, gw = T.shape
    nd = gw + gh - 1
    total = 0.0
    for d0 in range(0, nd, M):
        L = np.zeros_like(T)
        best = 0.0
        for d in range(d0, min(d0 + M, nd)):
            xs = np.arange(max(0, d - gh + 1), min(gw - 1, d) + 1)
            ys = d - xs
            up = np.where(ys[None, :] > 0, L[:, np.maximum(ys - 1, 0), xs], 0.0) if d > d0 else 0.0
            left = np.where(xs[None, :] > 0, L[:, ys, np.maximum(xs - 1, 0)], 0.0) if d > d0 else 0.0
            L[:, ys, xs] = T[:, ys, xs] + np.maximum(up, left)
            best = max(best, float(L[:, ys, xs].max()))
        total += best
    return total


def main():
    it = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    scene = bench.get_scene(2, 1.0)
    ctx = pmk.Context(nviews=scene.nviews, sweep_group=scene.nviews)
    ctx.set_scene(scene.P, scene.images)
    ctx.set_depth(0); ctx.store_clear(); ctx.store_add(*synth.seed_arrays(scene)); ctx.set_depth(1)
    gw, gh = ctx.grid_dims(0)
    for k in range(it):
        ctx.propagate(k, 0x5EED0001); ctx.filter(); ctx.update_threshold()
    pmk._chk(pmk.lib().pmk_debug_cell_times(ctx.h, None))
    st = ctx.propagate(it, 0x5EED0001)
    out = np.zeros(scene.nviews * gw * gh, np.float32)
    pmk._chk(pmk.lib().pmk_debug_cell_times(ctx.h, pmk._p(out)))
    T = out.reshape(scene.nviews, gh, gw).astype(np.float64) * 1e-9
    print(f"iteration {it}: cells {T.size}, busy cells {(T > 0).sum()}, total cell time {T.sum():.1f} s, slowest-cell sum (kernel) {st['step_max_ns'] / 1e9:.2f} s")
    for M in (1, 2, 4, 8, 16, 64, 10 ** 6):
        print(f"  barrier every {M:>7d} anti-diagonals: chain bound {chain(T, M):.3f} s")
    print(f"  work bound at 1184 resident warps: {T.sum() / 1184:.3f} s")


if __name__ == "__main__":
    main()
