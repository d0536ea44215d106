#!/usr/bin/env python
"""bench.py -- hypothesis-NCC throughput of the PM-MVS hot path on B200 (BASELINE.json metric).

One "step" = one pass of the hot path (K1: PatchManager::computeNcc-equivalent, optim.cpp:630-706) over one
batch of 2^20 synthetic hypotheses on the templeRing-shaped scene (BASELINE.json configs[1]: 47 views 640x480,
tau = 6 views per eval, 7x7 RGB lattice).  N > 1 shards hypotheses across ranks (no data-path collective:
the evals are independent; images are replicated), so scaling is weak.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # the product (CUDA, through the C ABI)
    python bench.py --impl reference ...                           # the reference's own CPU code, host cores

Prints ONE JSON line (rank 0).  `value` = whole-job evals/s with inputs resident in HBM (CUDA events on the
launching stream, L2 flushed between steps, max over ranks); `e2e` = the same metric through the host-buffer
C-ABI call (pinned host inputs, H2D and D2H inside the timed region).  `e2e_packed` (informational) = the same steps through
pmk_ncc_eval_packed, the byte-lean form of that call (31 B instead of 60 B per hypothesis over PCIe, bit-identical results).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_EVAL = 4692.0     # 6 views x 64 texels x 3 ch x 4 B + 84 B patch I/O
METRIC = "hypothesis_ncc_evals_per_sec"
UNIT = "evals/s"
WORKLOAD = "config2 templeRing-shaped synthetic scene: 47 views 640x480, 2^20 hypotheses/step/GPU around GT, tau=6, 7x7 RGB"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def load_traffic(n_per_launch: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed ncu capture."""
    path = os.path.join(ROOT, "profiles", "k1_traffic.json")
    try:
        with open(path) as fh:
            t = json.load(fh)
        per_hyp = (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["hypotheses_per_launch"]
        return per_hyp * n_per_launch, t["source"]
    except (OSError, KeyError, ValueError):
        return None, None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------------
def scene_cache_dir():
    d = os.path.join(tempfile.gettempdir(), "pmk_bench_cache")
    os.makedirs(d, exist_ok=True)
    return d


def get_scene(config: int, scale: float):
    """Render (or load the cached render of) the synthetic scene; deterministic, so every rank agrees."""
    from mvskit_b200 import synth
    scene = synth.make_scene(config, scale=scale)
    path = os.path.join(scene_cache_dir(), f"scene_c{config}_s{scale:g}.npy")
    if os.path.exists(path):
        arr = np.load(path)
        scene.images = [arr[v] for v in range(arr.shape[0])]
    else:
        scene.render()
        tmp = path + f".{os.getpid()}.tmp.npy"
        np.save(tmp, np.stack(scene.images))
        os.replace(tmp, path)
    return scene


def get_hypotheses(scene, n: int, seed: int, config: int, scale: float, order: str = "grid"):
    path = os.path.join(scene_cache_dir(), f"hyp_c{config}_s{scale:g}_n{n}_seed{seed}_{order}.npz")
    if os.path.exists(path):
        z = np.load(path)
        return z["c"], z["n"], z["v"], z["nv"]
    c, nrm, vw, nv = scene.hypotheses(n, seed=seed, order=order)
    tmp = path + f".{os.getpid()}.tmp.npz"
    np.savez(tmp, c=c, n=nrm, v=vw, nv=nv)
    os.replace(tmp, path)
    return c, nrm, vw, nv


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc:
            self.proc.terminate()
        sm, smax, reasons = [], 0.0, set()
        for t, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8 or not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            try:
                sm.append(float(parts[1])); smax = max(smax, float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# CPU arms (the only places bench.py executes anything under oracle/)
# ------------------------------------------------------------------------------------------------------
_REF = None


def _ref_init(prefix):
    global _REF
    from oracle.pyoracle import RefLib
    _REF = RefLib(prefix)


def _ref_worker(args):
    c, n, vw, nv = args
    return _REF.time_compute_ncc(c, n, vw, nv, 1)


class RefPool:
    """`cores` processes, each holding one instance of the compiled reference (it is not re-entrant:
    static Optim::m_inst, optim.cpp:18-22), timed on disjoint slices of the batch."""

    def __init__(self, prefix: str, cores: int):
        self.cores = cores
        if cores == 1:
            _ref_init(prefix)
            self.pool = None
        else:
            import multiprocessing as mp
            self.pool = mp.get_context("spawn").Pool(cores, initializer=_ref_init, initargs=(prefix,))

    def evals_per_sec(self, hyp, sample_per_core: int) -> float:
        c, n, vw, nv = (a[: sample_per_core * self.cores] for a in hyp)
        jobs = [(c[i::self.cores], n[i::self.cores], vw[i::self.cores], nv[i::self.cores]) for i in range(self.cores)]
        secs = [_ref_worker(jobs[0])] if self.pool is None else self.pool.map(_ref_worker, jobs)
        return len(c) / max(secs)

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def ensure_scene_dir(scene, config, scale):
    from mvskit_b200 import synth
    d = os.path.join(scene_cache_dir(), f"dir_c{config}_s{scale:g}")
    if not os.path.exists(os.path.join(d, "option")):
        tmp = d + f".{os.getpid()}.tmp"
        synth.write_scene(scene, tmp, with_seeds=False)
        try:
            os.replace(tmp, d)
        except OSError:
            pass
    return d + "/"


def cpu_reference_evals_per_sec(scene, config, scale, hyp, cores: int, sample: int, steps: int = 1, warmup: int = 0):
    """Time the reference's own computeNcc loop (oracle/_ref/libpmref.so) on `cores` processes, `sample` evals each."""
    from oracle import pyoracle
    if os.path.exists(pyoracle.REF_SO):
        pool = RefPool(ensure_scene_dir(scene, config, scale), cores)
        vals = [pool.evals_per_sec(hyp, sample) for _ in range(warmup + steps)][warmup:]
        pool.close()
        return float(np.mean(vals)), "reference"
    pyoracle.build(ref=False)
    orc = pyoracle.COracle(scene.P, scene.images)
    c, n, vw, nv = (a[:sample] for a in hyp)
    t0 = time.time()
    orc.compute_ncc(c, n, vw, nv)
    return sample / (time.time() - t0), "port"


# ------------------------------------------------------------------------------------------------------
# second BASELINE metric: end-to-end patches/sec of init -> (propagate, filter, updateThreshold) x ITER
# ------------------------------------------------------------------------------------------------------
def run_pipeline(ctx, scene, iters: int, seed: int = 0x5EED0001):
    """PmMvps::run (pmmvps.cpp:76-114) through the C ABI on the scene already resident in `ctx`: seeds -> patch store,
    then ITER x (Propagate::run, Filter::run, updateThreshold).  Wall time includes the seed upload and every host sync."""
    from mvskit_b200 import synth
    coord, normal, scal, images, nimg = synth.seed_arrays(scene)
    l0 = ctx.launch_count()
    ctx.sync()
    t0 = time.perf_counter()
    ctx.set_depth(0)
    ctx.store_clear()
    ctx.store_add(coord, normal, scal, images, nimg)
    ctx.set_depth(1)
    evals, calls, t_prop, t_filt, counts = 0, 0, 0.0, 0.0, None
    for it in range(iters):
        t = time.perf_counter()
        st = ctx.propagate(it, seed)
        ctx.sync()
        t_prop += time.perf_counter() - t
        evals += st["evals"]
        calls += st["calls"]
        t = time.perf_counter()
        counts = ctx.filter()
        ctx.sync()
        t_filt += time.perf_counter() - t
        ctx.update_threshold()
    n = ctx.store_count()
    dt = time.perf_counter() - t0
    return {"patches": n, "seconds": dt, "patches_per_sec": n / dt, "seeds": int(len(coord)), "iters": iters, "propagate_seconds": t_prop,
            "propagate_patch_calls": int(calls), "propagate_patch_calls_per_sec": calls / max(t_prop, 1e-9),
            "filter_seconds": t_filt, "sweep_ncc_evals": int(evals), "sweep_ncc_evals_per_sec": evals / max(t_prop, 1e-9),
            "last_filter_counts": counts, "gpu_launches": ctx.launch_count() - l0}


def cpu_reference_sweep_calls_per_sec(scene, config, scale, budget_s: float = 15.0):
    """The reference's own Propagate::propagatePatch (oracle/_ref/libpmref.so), driven dest cell by dest cell over the first
    anti-diagonals of view 0 from the same seeds, for `budget_s` seconds on one core: propagatePatch calls per second."""
    from oracle import pyoracle
    from mvskit_b200 import synth
    if not os.path.exists(pyoracle.REF_SO):
        return None
    global _REF
    if _REF is None:
        _ref_init(ensure_scene_dir(scene, config, scale))
    ref = _REF
    coord, normal, scal, images, nimg = synth.seed_arrays(scene)
    ref.clear_patches()
    ref.set_depth(0)
    ref.add_patches(coord, normal, scal, images, nimg)
    ref.set_depth(1)
    ref.refine_seed(0x5EED0001)
    gw, gh = ref.grid_dims(0)
    calls, d, t0 = 0, 12, time.perf_counter()
    while time.perf_counter() - t0 < budget_s and d < gw + gh - 1:
        calls += ref.propagate_diag(0, d, 1, 0)
        d += 1
    dt = time.perf_counter() - t0
    return {"value": calls / dt, "unit": "propagatePatch calls/s", "cores": 1, "kind": "reference",
            "sample": f"{calls} calls: anti-diagonals 12..{d - 1} of view 0, iteration 0, same seeds, {dt:.1f} s"}


# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="pmk", choices=["pmk", "reference"])
    ap.add_argument("--batch", type=int, default=1 << 20, help="hypotheses per step per GPU")
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--cpu-sample", type=int, default=1 << 20, help="evals in the cpu_baseline sample (1 core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipeline-iters", type=int, default=3, help="ITER of the end-to-end patches/sec run (0 = skip)")
    ap.add_argument("--sweep-group", type=int, default=0, help="views swept together in Propagate::run (0 = all)")
    ap.add_argument("--order", default="grid", choices=["grid", "random"], help="hypothesis order: Z-order of the reference pixel (patch-grid walk) or random")
    args = ap.parse_args()

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    n_gpus = max(args.gpus, world)
    warmup = max(args.warmup, 3) if args.impl == "pmk" else args.warmup
    config = {"workload": WORKLOAD if (args.config == 2 and args.scale == 1.0 and args.batch == 1 << 20) else
              f"config{args.config} scale {args.scale:g}, {args.batch} hypotheses/step/GPU", "views_per_eval": 6,
              "hypotheses_per_step_per_gpu": args.batch, "sharding": "hypotheses split across ranks, images replicated, no collective",
              "hypothesis_order": args.order, "l2": "flushed (256 MiB memset) before every timed step"}

    # ---------------- reference arm: the reference's own CPU code on the host cores ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        scene = get_scene(args.config, args.scale)
        cores = os.cpu_count() or 1
        per_core = max(1, min(args.batch // cores, 32768))          # bounded sample per step (~0.3 s)
        hyp = get_hypotheses(scene, args.batch, 7, args.config, args.scale, args.order)
        value, kind = cpu_reference_evals_per_sec(scene, args.config, args.scale, hyp, cores, per_core, args.steps, args.warmup)
        sample = f"{per_core * cores} of the {args.batch} hypotheses per step ({per_core} per process x {cores} processes)"
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * per_core * cores / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ---------------- product arm ----------------
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod
    from mvskit_b200 import pmk

    if local_rank == 0 or world == 1:
        scene = get_scene(args.config, args.scale)
    if dist:
        dist.barrier()
        if local_rank != 0:
            scene = get_scene(args.config, args.scale)
    hyp = get_hypotheses(scene, args.batch, 7 + rank, args.config, args.scale, args.order)
    c, n, vw, nv = hyp
    N = len(c)

    ctx = pmk.Context(nviews=scene.nviews, device=local_rank, sweep_group=args.sweep_group or scene.nviews)
    ctx.set_scene(scene.P, scene.images)
    launches0 = ctx.launch_count()

    # device-resident inputs for `value`
    d_c, d_n, d_v, d_nv = ctx.alloc(c.nbytes).upload(c), ctx.alloc(n.nbytes).upload(n), ctx.alloc(vw.nbytes).upload(vw), ctx.alloc(nv.nbytes).upload(nv)
    d_incc, d_ncc = ctx.alloc(N * 4), ctx.alloc(N * 4)
    # pinned host inputs/outputs for `e2e`
    h_c, h_n, h_v, h_nv = (pmk.pinned_empty(a.shape, a.dtype) for a in (c, n, vw, nv))
    h_c[:], h_n[:], h_v[:], h_nv[:] = c, n, vw, nv
    h_incc, h_ncc = pmk.pinned_empty((N,), np.float32), pmk.pinned_empty((N,), np.float32)
    ctx.sync()

    def barrier():
        ctx.sync()
        if dist:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    def step_dev():
        ctx.ncc_eval_dev(N, d_c, d_n, d_v, d_nv, vw.shape[1], d_incc, d_ncc)

    def step_e2e():
        pmk._chk(pmk.lib().pmk_ncc_eval(ctx.h, N, pmk._p(h_c), pmk._p(h_n), pmk._p(h_v), pmk._p(h_nv), vw.shape[1], pmk._p(h_incc), pmk._p(h_ncc), None))

    for _ in range(warmup):
        step_dev()
    step_e2e()
    # workload statistics (untimed): how many of the tau views of each eval actually get sampled
    lv = ctx.ncc_eval(c[: 1 << 16], n[: 1 << 16], vw[: 1 << 16], nv[: 1 << 16], want_levels=True)[2]
    valid_views = float((lv >= 0).sum(1).mean())
    launches_extra = 1
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    # ---- timed: K device-resident steps ----
    barrier()
    t_wall0 = time.time()
    ms_steps = []
    for _ in range(args.steps):
        ctx.flush_l2()
        ctx.timer_begin()
        step_dev()
        ms_steps.append(ctx.timer_end())
    barrier()
    t_wall1 = time.time()
    launches_timed = ctx.launch_count() - launches0 - warmup - 1 - launches_extra
    # ---- timed: K end-to-end steps (host buffers, copies inside) ----
    e2e_ms = []
    for _ in range(args.steps):
        ctx.flush_l2()
        ctx.sync()
        t0 = time.perf_counter()
        step_e2e()                      # returns after the D2H landed
        e2e_ms.append(1e3 * (time.perf_counter() - t0))
    # ---- the same end-to-end steps through the byte-lean call (pmk_ncc_eval_packed: 31 B in per eval instead of 60); reported beside
    # `e2e` as `e2e_packed`, never in its place: the headline call mirrors the reference's Patch fields one for one ----
    packed_ms = []
    if (c[:, 3] == 1).all() and scene.nviews <= 255:
        pc, pn_, pv, pnv = ctx.pack_hypotheses(c, n, vw, nv)
        hp = [pmk.pinned_empty(a.shape, a.dtype) for a in (pc, pn_, pv, pnv)]
        for dst, src in zip(hp, (pc, pn_, pv, pnv)):
            dst[:] = src
        packed_bytes = int(sum(a.nbytes for a in hp))

        def step_packed():
            pmk._chk(pmk.lib().pmk_ncc_eval_packed(ctx.h, N, pmk._p(hp[0]), pmk._p(hp[1]), pmk._p(hp[2]), pmk._p(hp[3]), vw.shape[1],
                                                   pmk._p(h_incc), pmk._p(h_ncc), None))
        step_packed()
        for _ in range(min(args.steps, 50)):
            ctx.flush_l2()
            ctx.sync()
            t0 = time.perf_counter()
            step_packed()
            packed_ms.append(1e3 * (time.perf_counter() - t0))
    barrier()
    clocks = sampler.stop(t_wall0, time.time())

    # BASELINE metric (2) on N > 1 GPUs: the same scene, dest-cell rows split into bands, step mutations all-gathered (NCCL)
    pipe_mg = None
    if dist and args.pipeline_iters > 0:
        import torch
        from mvskit_b200 import dist as pdist
        try:
            pdist.init_comm(ctx, dist, device=torch.device("cuda", local_rank))
            barrier()
            pipe_mg = run_pipeline(ctx, scene, args.pipeline_iters)
            barrier()
            digest = ctx.store_checksum()
            allsum = [None] * world
            dist.all_gather_object(allsum, digest)
            secs = torch.tensor([pipe_mg["seconds"]], device="cuda", dtype=torch.float64)
            dist.all_reduce(secs, op=dist.ReduceOp.MAX)
            pipe_mg["seconds"] = float(secs[0])
            pipe_mg["patches_per_sec"] = pipe_mg["patches"] / pipe_mg["seconds"]
            pipe_mg["replicas_identical"] = all(x == allsum[0] for x in allsum)
            pipe_mg["config"] = (f"config{args.config} scale {args.scale:g}: {scene.nviews} views, {world} GPUs (row bands, replicated store, "
                                 f"ncclAllGather of step mutations), sweep_group {args.sweep_group or scene.nviews}")
        except Exception as exc:
            pipe_mg = {"error": str(exc)}

    total_ms, total_e2e, packed_total = float(sum(ms_steps)), float(sum(e2e_ms)), float(sum(packed_ms))
    if dist:
        import torch
        t = torch.tensor([total_ms, total_e2e, packed_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, total_e2e, packed_total = float(t[0]), float(t[1]), float(t[2])
    evals = float(N) * args.steps * world
    value = evals / (total_ms * 1e-3)
    e2e_value = evals / (total_e2e * 1e-3)

    if rank == 0:
        peak, peak_src = load_peaks()
        per_gpu = value / world
        # algorithmic bytes per eval = SURVEY.md 8(d)'s figure: 84 B of patch I/O + 768 B (64 texels x 3 ch x 4 B) per
        # sampled view (4692 B at tau = 6); views failing the angle gate / getTexSafe are not gathered, so the texel
        # term is scaled by the measured valid-view count.  The pyramid is STORED as 4 x fp16 texels (exact for the
        # u8-rounded values), so the bytes this layout must move are 84 + 384 per view: reported as *_layout.
        algo_bytes = 84.0 + 768.0 * valid_views
        layout_bytes = 84.0 + 384.0 * valid_views
        achieved = per_gpu * algo_bytes / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(launches_timed),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(c.nbytes + n.nbytes + vw.nbytes + nv.nbytes),
                    "d2h_bytes_per_step": int(2 * N * 4), "ms_per_step": total_e2e / args.steps},
            "e2e_packed": ({"value": float(N) * len(packed_ms) * world / (packed_total * 1e-3), "unit": UNIT, "h2d_bytes_per_step": packed_bytes,
                            "d2h_bytes_per_step": int(2 * N * 4), "ms_per_step": packed_total / len(packed_ms), "steps": len(packed_ms),
                            "call": "pmk_ncc_eval_packed (3-float coord / normal, byte view ids); informational, `e2e` is the headline"}
                           if packed_ms else None),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": load_traffic(N)[0],
                         "traffic_unit": "bytes/launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)", "traffic_source": load_traffic(N)[1],
                         "algorithmic_bytes_per_launch": algo_bytes * N,
                         "kernel": "k1_ncc<7,4>", "algorithmic_bytes_per_eval": algo_bytes, "valid_views_per_eval": valid_views,
                         "layout_bytes_per_eval": layout_bytes, "achieved_layout": per_gpu * layout_bytes / 1e9, "frac_layout": per_gpu * layout_bytes / 1e9 / peak,
                         "note": "not HBM-bound: the pyramid stays L2/L1 resident; binding units per ncu (profiles/r01_k1_ncc_v4_streamed.txt): L1TEX data pipe 85%, issue slots 78%", "peak_source": peak_src,
                         "wall_ms_timed_region": 1e3 * (t_wall1 - t_wall0)},
        }
        if pipe_mg is not None:
            out["pipeline"] = pipe_mg
        if world == 1 and args.pipeline_iters > 0:
            # BASELINE metric (2): patches alive after the last Filter::run / wall time of init -> propagate x ITER -> filter
            try:
                pipe = run_pipeline(ctx, scene, args.pipeline_iters)
            except Exception as exc:                      # the headline line must not be lost to the second metric
                pipe = {"error": str(exc)}
            pipe["config"] = f"config{args.config} scale {args.scale:g}: {scene.nviews} views, seeds every 4th cell, sweep_group {args.sweep_group or scene.nviews}"
            if "error" not in pipe and not args.no_cpu_baseline:
                try:
                    pipe["cpu_baseline"] = cpu_reference_sweep_calls_per_sec(scene, args.config, args.scale)
                except Exception as exc:
                    pipe["cpu_baseline"] = {"error": str(exc)}
            out["pipeline"] = pipe
        if world == 1 and not args.no_cpu_baseline:
            v, kind = cpu_reference_evals_per_sec(scene, args.config, args.scale, hyp, 1, min(args.cpu_sample, N))
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": kind,
                                   "sample": f"first {min(args.cpu_sample, N)} hypotheses of the step's batch, PatchManager::computeNcc loop, 1 process"}
        print(json.dumps(out))
    ctx.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
