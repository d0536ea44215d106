#!/usr/bin/env python
"""Stress of the streamed host-buffer NCC call (pmk_ncc_eval): many calls of random sizes through pinned and pageable buffers, each
compared bit for bit with the device-pointer call on the same inputs.  A lost arrival word would show as a hang (run under `timeout`),
a stale one as a mismatch.   usage: python tools/stress_stream.py [calls=1500]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvskit_b200 import pmk, synth  # noqa: E402


def main():
    calls = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    scene = synth.make_scene(1, scale=0.5).render()
    ctx = pmk.Context(nviews=scene.nviews)
    ctx.set_scene(scene.P, scene.images)
    N = (1 << 19) + 4321
    c, n, vw, nv = scene.hypotheses(N, seed=5, well_observed=False)
    d = [ctx.alloc(a.nbytes).upload(a) for a in (c, n, vw, nv)]
    d_incc, d_ncc = ctx.alloc(N * 4), ctx.alloc(N * 4)
    ctx.ncc_eval_dev(N, d[0], d[1], d[2], d[3], vw.shape[1], d_incc, d_ncc)
    want_incc, want_ncc = np.empty(N, np.float32), np.empty(N, np.float32)
    d_incc.download(want_incc); d_ncc.download(want_ncc)
    h = [pmk.pinned_empty(a.shape, a.dtype) for a in (c, n, vw, nv)]
    for dst, src in zip(h, (c, n, vw, nv)):
        dst[:] = src
    h_incc, h_ncc = pmk.pinned_empty((N,), np.float32), pmk.pinned_empty((N,), np.float32)
    rng = np.random.default_rng(11)
    t0 = time.time()
    total = 0
    for k in range(calls):
        m = int(rng.choice([rng.integers(1, 70000), rng.integers(1, N + 1), (1 << 14) * int(rng.integers(1, 33)) + int(rng.integers(-1, 2))]))
        m = max(1, min(m, N))
        o = int(rng.integers(0, N - m + 1))          # a window of the inputs: every call sees different data at the same slots
        if k % 5 == 4:                               # pageable buffers
            incc, ncc = ctx.ncc_eval(c[o:o + m], n[o:o + m], vw[o:o + m], nv[o:o + m])
        else:
            h_incc[:m] = -3.0
            pmk._chk(pmk.lib().pmk_ncc_eval(ctx.h, m, pmk._p(h[0][o:o + m]), pmk._p(h[1][o:o + m]), pmk._p(h[2][o:o + m]), pmk._p(h[3][o:o + m]),
                                            vw.shape[1], pmk._p(h_incc), pmk._p(h_ncc), None))
            incc, ncc = h_incc[:m], h_ncc[:m]
        ok = np.array_equal(incc.view(np.uint32), want_incc[o:o + m].view(np.uint32)) and np.array_equal(ncc.view(np.uint32), want_ncc[o:o + m].view(np.uint32))
        if not ok:
            bad = np.nonzero(incc.view(np.uint32) != want_incc[o:o + m].view(np.uint32))[0]
            print(f"MISMATCH call {k}: n={m} offset={o} first bad {bad[:5]} of {len(bad)}")
            sys.exit(1)
        total += m
    print(f"stress ok: {calls} calls, {total} evals, {time.time() - t0:.1f} s, all bit-identical to the device-pointer call")
    ctx.close()


if __name__ == "__main__":
    main()
