#!/usr/bin/env python
"""Where does a propagatePatch try spend its time?  Runs the config-2 pipeline (all views per step) with the sweep's phase recording on
(pmk_debug_phase_times) and prints, per iteration, the warp time of each phase of a try summed over all warps, its share, and the mean
per try / per refined try.  Profiling only: the recording adds one global atomic per phase boundary.
usage: python tools/phase_times.py [iterations=2]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mvskit_b200 import pmk, synth  # noqa: E402

NAMES = ("generatePatch + computeNcc", "preProcess", "refinePatch (PMR1)", "postProcess (no store)", "postProcess tail (vimages, check)", "wait for turn / redo")


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    scene = bench.get_scene(2, 1.0)
    ctx = pmk.Context(nviews=scene.nviews, sweep_group=scene.nviews)
    ctx.set_scene(scene.P, scene.images)
    ctx.set_depth(0); ctx.store_clear(); ctx.store_add(*synth.seed_arrays(scene)); ctx.set_depth(1)
    pmk._chk(pmk.lib().pmk_debug_phase_times(ctx.h, None))
    for it in range(iters):
        t = time.perf_counter()
        st = ctx.propagate(it, 0x5EED0001)
        ctx.sync()
        wall = time.perf_counter() - t
        ph = np.zeros(16, np.uint64)
        pmk._chk(pmk.lib().pmk_debug_phase_times(ctx.h, pmk._p(ph)))
        tot = float(ph[:6].sum())
        tries, refined = int(ph[6]), int(ph[7])
        print(f"iteration {it}: propagate {wall:.2f} s wall, {tries} tries ({refined} refined), cell time {st['cell_ns'] / 1e9:.1f} s, "
              f"slowest-cell sum {st['step_max_ns'] / 1e9:.2f} s, phase time {tot / 1e9:.1f} s")
        for i, name in enumerate(NAMES):
            per = refined if i in (2, 3, 4) else tries
            print(f"  {name:<36s} {float(ph[i]) / 1e9:8.2f} s  {100.0 * float(ph[i]) / max(tot, 1):5.1f} %   {float(ph[i]) / max(per, 1) / 1e3:8.1f} us per {'refined try' if i in (2, 3, 4) else 'try'}")
        if ph[8:].any():
            sub = ph[8:].astype(float) / 1e9
            print("  cta_costs sub-phases (s): candidates %.2f, frames %.2f, sampling %.2f, dots %.2f, argmin %.2f" % tuple(sub[:5]))
        ctx.filter(); ctx.update_threshold()
    ctx.close()


if __name__ == "__main__":
    main()
