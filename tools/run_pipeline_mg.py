#!/usr/bin/env python
"""Multi-GPU end-to-end run (one process per GPU under torchrun): replicated store, the dest cells of every step dealt out to the ranks in turn, the step
mutations all-gathered over NCCL.  Checks that every replica ends with the same store and (rank 0) that the result equals a
single-GPU run of the same scene.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/run_pipeline_mg.py --config 1 --scale 0.5"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from mvskit_b200 import dist as pdist, pmk, synth  # noqa: E402


def pipeline(ctx, seeds, iters, seed):
    coord, normal, scal, images, nimg = seeds
    ctx.set_depth(0); ctx.store_clear(); ctx.store_add(coord, normal, scal, images, nimg); ctx.set_depth(1)
    sums = []
    evals = 0
    for it in range(iters):
        evals += ctx.propagate(it, seed)["evals"]
        sums.append(ctx.store_checksum())
        ctx.filter()
        sums.append(ctx.store_checksum())
        ctx.update_threshold()
    return sums, evals


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--scale", type=float, default=0.5)
    ap.add_argument("--nviews", type=int, default=None)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--group", type=int, default=0)
    ap.add_argument("--no-single", action="store_true")
    ap.add_argument("--procs", type=int, default=1)
    ap.add_argument("--seed-stride", type=int, default=4)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    # rank 0 renders (in parallel processes) and caches; the other ranks load the cache
    import tempfile
    from run_pipeline import render_parallel
    cache = os.path.join(tempfile.gettempdir(), f"pmk_mg_scene_c{a.config}_s{a.scale:g}_v{a.nviews}.npy")
    scene = synth.make_scene(a.config, scale=a.scale, nviews=a.nviews)
    if rank == 0 and not os.path.exists(cache):
        render_parallel(scene, a.config, a.scale, a.nviews, a.procs) if a.procs > 1 else scene.render()
        np.save(cache + ".tmp.npy", np.stack(scene.images))
        os.replace(cache + ".tmp.npy", cache)
    dist.barrier()
    arr = np.load(cache)
    scene.images = [arr[v] for v in range(arr.shape[0])]
    seeds = synth.seed_arrays(scene, stride=a.seed_stride)
    group = a.group or scene.nviews
    ctx = pmk.Context(nviews=scene.nviews, device=local, sweep_group=group)
    ctx.set_scene(scene.P, scene.images)
    pdist.init_comm(ctx, dist, device=torch.device("cuda", local))
    dist.barrier()
    t0 = time.perf_counter()
    sums, evals = pipeline(ctx, seeds, a.iters, 0x5EED0001)
    ctx.sync()
    torch.cuda.synchronize()
    dist.barrier()
    dt = time.perf_counter() - t0
    allsums = [None] * world
    dist.all_gather_object(allsums, sums)
    n = ctx.store_count()
    ok_replicas = all(s == allsums[0] for s in allsums)
    out = {"n_gpus": world, "patches": n, "seconds": dt, "patches_per_sec": n / dt, "replicas_identical": ok_replicas, "rank_evals": evals}
    if rank == 0 and not a.no_single:
        one = pmk.Context(nviews=scene.nviews, device=local, sweep_group=group)
        one.set_scene(scene.P, scene.images)
        t0 = time.perf_counter()
        s1, e1 = pipeline(one, seeds, a.iters, 0x5EED0001)
        one.sync()
        out["single_gpu_seconds"] = time.perf_counter() - t0
        out["equals_single_gpu"] = s1 == sums
        out["single_gpu_patches"] = one.store_count()
        if s1 != sums:
            out["first_diff"] = next(i for i, (x, y) in enumerate(zip(s1, sums)) if x != y)
            out["sums_multi"], out["sums_single"] = sums[:4], s1[:4]
        one.close()
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
