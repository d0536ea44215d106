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


def host_procs(world: int = 1) -> int:
    return max(1, (os.cpu_count() or 1) // max(world, 1))


def get_scene(config: int, scale: float, rank: int = 0, world: int = 1, barrier=None):
    """Render (or load the cached render of) the synthetic scene.  Deterministic, so every rank agrees; with several ranks each one
    renders its share of the views with its share of the host cores into a per-view cache, then all load all views."""
    from mvskit_b200 import synth
    scene = synth.make_scene(config, scale=scale)
    d = os.path.join(scene_cache_dir(), f"scene_c{config}_s{scale:g}")
    os.makedirs(d, exist_ok=True)
    V = scene.nviews
    mine = [v for v in list(range(V))[rank::world] if not os.path.exists(os.path.join(d, f"view{v:03d}.npy"))]
    if mine:
        for v, im in synth.render_views(config, scale, mine, host_procs(world)):
            tmp = os.path.join(d, f"view{v:03d}.{os.getpid()}.tmp.npy")
            np.save(tmp, im)
            os.replace(tmp, os.path.join(d, f"view{v:03d}.npy"))
    if barrier:
        barrier()
    scene.images = [np.load(os.path.join(d, f"view{v:03d}.npy")) for v in range(V)]
    return scene


def get_seeds(scene, config: int, scale: float, stride: int, rank: int = 0, world: int = 1, barrier=None):
    """Seed patch arrays (coord, normal, scal, images, nimages).  Configs 1-2: the single-stream generator used since round 1;
    bigger configs: one stream per view, generated in parallel over ranks and host processes, cached per view group."""
    from mvskit_b200 import synth
    if config <= 2:
        return synth.seed_arrays(scene, stride=stride)
    d = os.path.join(scene_cache_dir(), f"seeds_c{config}_s{scale:g}_st{stride}")
    os.makedirs(d, exist_ok=True)
    V = scene.nviews
    path = os.path.join(d, f"part{rank}of{world}.npz")
    if not os.path.exists(path):
        recs = synth.seed_records(config, scale, list(range(V))[rank::world], host_procs(world), stride=stride)
        arrs = synth.seed_arrays(scene, recs=recs)
        tmp = path + f".{os.getpid()}.tmp.npz"
        np.savez(tmp, *arrs)
        os.replace(tmp, path)
    if barrier:
        barrier()
    parts = []
    for r in range(world):
        z = np.load(os.path.join(d, f"part{r}of{world}.npz"))
        parts.append([z[f"arr_{i}"] for i in range(5)])
    arrs = [np.concatenate([p[i] for p in parts]) for i in range(5)]
    order = np.argsort(arrs[3][:, 0], kind="stable")                 # ascending reference view, as one process would produce them
    return tuple(a[order] for a in arrs)


def _hyp_job(args):
    config, scale, n, seed, order = args
    from mvskit_b200 import synth
    return synth.make_scene(config, scale=scale).hypotheses(n, seed=seed, order=order)


def get_hypotheses(scene, n: int, seed: int, config: int, scale: float, order: str = "grid", procs: int = 1):
    path = os.path.join(scene_cache_dir(), f"hyp_c{config}_s{scale:g}_n{n}_seed{seed}_{order}_p{procs}.npz")
    if os.path.exists(path):
        z = np.load(path)
        return z["c"], z["n"], z["v"], z["nv"]
    if procs <= 1:
        c, nrm, vw, nv = scene.hypotheses(n, seed=seed, order=order)
    else:                                                             # big configs: visibility over 49-128 views is the cost; split by seed
        from mvskit_b200 import synth
        per = (n + procs - 1) // procs
        parts = synth._pool_map(_hyp_job, [(config, scale, per, seed * 1000 + i, order) for i in range(procs)], procs)
        c, nrm, vw, nv = (np.concatenate([p[i] for p in parts])[:n] for i in range(4))
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
def run_pipeline(ctx, seeds, iters: int, seed: int = 0x5EED0001):
    """PmMvps::run (pmmvps.cpp:76-114) through the C ABI on the scene already resident in `ctx`: seeds -> patch store,
    then ITER x (Propagate::run, Filter::run, updateThreshold).  Wall time includes the seed upload and every host sync."""
    coord, normal, scal, images, nimg = seeds
    l0 = ctx.launch_count()
    ctx.sync()
    t0 = time.perf_counter()
    ctx.set_depth(0)
    ctx.store_clear()
    ctx.store_add(coord, normal, scal, images, nimg)
    ctx.set_depth(1)
    evals, calls, t_prop, t_filt, counts, nccl_ns, msg_bytes, steps = 0, 0, 0.0, 0.0, None, 0, 0, 0
    per_iter = []
    for it in range(iters):
        t = time.perf_counter()
        st = ctx.propagate(it, seed)
        ctx.sync()
        tp = time.perf_counter() - t
        t_prop += tp
        evals += st["evals"]
        calls += st["calls"]
        nccl_ns += st.get("nccl_ns", 0)
        msg_bytes += st.get("msg_bytes", 0)
        steps += st["steps"]
        t = time.perf_counter()
        counts = ctx.filter()
        ctx.sync()
        tf = time.perf_counter() - t
        t_filt += tf
        per_iter.append({"propagate_seconds": tp, "filter_seconds": tf, "patches": int(counts[5]), "calls": int(st["calls"]),
                         "slowest_cell_sum_seconds": st["step_max_ns"] / 1e9, "cell_seconds": st["cell_ns"] / 1e9})
        ctx.update_threshold()
    n = ctx.store_count()
    dt = time.perf_counter() - t0
    return {"patches": n, "seconds": dt, "patches_per_sec": n / dt, "seeds": int(len(coord)), "iters": iters, "propagate_seconds": t_prop,
            "propagate_patch_calls": int(calls), "propagate_patch_calls_per_sec": calls / max(t_prop, 1e-9),
            "filter_seconds": t_filt, "sweep_ncc_evals": int(evals), "sweep_ncc_evals_per_sec": evals / max(t_prop, 1e-9),
            "wavefront_steps": int(steps), "allgather_bytes_per_step": (msg_bytes // steps) if steps else 0, "nccl_seconds": nccl_ns / 1e9,
            "last_filter_counts": counts, "per_iteration": per_iter, "gpu_launches": ctx.launch_count() - l0}


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
    ap.add_argument("--pipelines", default="2:3:4,3:4:8,5:1:8", help="config:ITER:seed-stride of the end-to-end runs (the --config entry uses --pipeline-iters)")
    ap.add_argument("--pipeline-budget", type=float, default=240.0, help="seconds after which no further extra pipeline is started")
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

    scene = get_scene(args.config, args.scale, rank, world, (dist.barrier if dist else None))
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

    total_ms, total_e2e, packed_total = float(sum(ms_steps)), float(sum(e2e_ms)), float(sum(packed_ms))
    if dist:
        import torch
        t = torch.tensor([total_ms, total_e2e, packed_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, total_e2e, packed_total = float(t[0]), float(t[1]), float(t[2])
    evals = float(N) * args.steps * world
    value = evals / (total_ms * 1e-3)
    e2e_value = evals / (total_e2e * 1e-3)

    # ---------------- BASELINE metric (2): end-to-end patches/sec, on the configs the north-star names ----------------
    # config 2 (the K1 scene, already resident), then config 3 (DTU-shaped, 49 x 1600x1200, ITER 4) and config 5 (128 x 1920x1080,
    # ITER 1), each at its stated shape; N > 1: every step's dest cells dealt out to the ranks in turn, step mutations gathered (NCCL).
    t_bench0 = time.time()
    pipelines, k1_other = [], []
    plan = []
    for item in args.pipelines.split(","):
        if item.strip():
            cfg_i, it_i, st_i = (int(x) for x in item.split(":"))
            plan.append((cfg_i, it_i, st_i))
    if args.pipeline_iters <= 0:
        plan = []
    for cfg_i, it_i, st_i in plan:
        if cfg_i == args.config:
            it_i = args.pipeline_iters
        entry = {"config_id": cfg_i, "iters": it_i, "seed_stride": st_i}
        if time.time() - t_bench0 > args.pipeline_budget and cfg_i != args.config:
            entry["skipped"] = f"time budget of {args.pipeline_budget:.0f} s for the extra pipelines used up"
            pipelines.append(entry)
            continue
        try:
            t_prep = time.time()
            if cfg_i == args.config:
                ctx_i, scene_i, scale_i = ctx, scene, args.scale
            else:
                scale_i = args.scale
                scene_i = get_scene(cfg_i, scale_i, rank, world, (dist.barrier if dist else None))
                ctx_i = pmk.Context(nviews=scene_i.nviews, device=local_rank, sweep_group=args.sweep_group or scene_i.nviews)
                ctx_i.set_scene(scene_i.P, scene_i.images)
            seeds_i = get_seeds(scene_i, cfg_i, scale_i, st_i, rank, world, (dist.barrier if dist else None))
            entry["prepare_seconds"] = time.time() - t_prep
            if dist:
                import torch
                from mvskit_b200 import dist as pdist
                pdist.init_comm(ctx_i, dist, device=torch.device("cuda", local_rank))
            barrier()
            res = run_pipeline(ctx_i, seeds_i, it_i)
            barrier()
            if dist:
                import torch
                digest = ctx_i.store_checksum()
                allsum = [None] * world
                dist.all_gather_object(allsum, digest)
                t = torch.tensor([res["seconds"], res["propagate_seconds"], res["filter_seconds"], res["nccl_seconds"]], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                res["seconds"], res["propagate_seconds"], res["filter_seconds"], res["nccl_seconds"] = (float(x) for x in t)
                calls = torch.tensor([float(res["propagate_patch_calls"])], device="cuda", dtype=torch.float64)
                allcalls = [torch.zeros_like(calls) for _ in range(world)]
                dist.all_gather(allcalls, calls)
                res["propagate_patch_calls_per_rank"] = [int(x.item()) for x in allcalls]
                res["propagate_patch_calls"] = int(sum(res["propagate_patch_calls_per_rank"]))
                res["patches_per_sec"] = res["patches"] / res["seconds"]
                res["replicas_identical"] = all(x == allsum[0] for x in allsum)
                res["store_checksum"] = [int(x) for x in digest]
            else:
                res["store_checksum"] = [int(x) for x in ctx_i.store_checksum()]
            entry.update(res)
            entry["config"] = (f"config{cfg_i} scale {scale_i:g}: {scene_i.nviews} views {scene_i.width}x{scene_i.height}, seeds every {st_i}th cell, "
                               f"ITER {it_i}, sweep_group {args.sweep_group or scene_i.nviews}, {world} GPU(s)"
                               + (": every step's dest cells dealt out to the ranks in turn, replicated store, NCCL all-gather of each step's mutations" if world > 1 else ""))
            if cfg_i != args.config:
                # K1 on this config's pyramid (not L2-resident): device-resident steps, L2 flushed, same kernel
                nh = 1 << 19                                  # enough batches for every resident warp (a 2^17 launch is all tail)
                hyp_i = get_hypotheses(scene_i, nh, 7 + rank, cfg_i, scale_i, args.order, procs=host_procs(world))
                ci, ni, vi, nvi = hyp_i
                bufs = [ctx_i.alloc(a.nbytes).upload(a) for a in (ci, ni, vi, nvi)]
                oi, on = ctx_i.alloc(len(ci) * 4), ctx_i.alloc(len(ci) * 4)
                for _ in range(3):
                    ctx_i.ncc_eval_dev(len(ci), bufs[0], bufs[1], bufs[2], bufs[3], vi.shape[1], oi, on)
                ms_i = []
                for _ in range(10):
                    ctx_i.flush_l2()
                    ctx_i.timer_begin()
                    ctx_i.ncc_eval_dev(len(ci), bufs[0], bufs[1], bufs[2], bufs[3], vi.shape[1], oi, on)
                    ms_i.append(ctx_i.timer_end())
                pyr = sum(int(scene_i.width >> l) * int(scene_i.height >> l) * 8 for l in range(4)) * scene_i.nviews
                k1_other.append({"config_id": cfg_i, "views": scene_i.nviews, "image": f"{scene_i.width}x{scene_i.height}", "pyramid_bytes": pyr,
                                 "hypotheses_per_step": len(ci), "steps": 10, "ms_per_step": float(np.mean(ms_i)), "l2_prefetch": os.environ.get("PMK_K1_PREFETCH", "auto (on: sampled levels exceed 3/4 of L2)"),
                                 "value_per_gpu": len(ci) / (float(np.mean(ms_i)) * 1e-3), "unit": UNIT})
                ctx_i.close()
        except Exception as exc:                      # the headline line must not be lost to the second metric
            entry["error"] = f"{type(exc).__name__}: {exc}"
        pipelines.append(entry)

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
            "roofline": {"bound": "l1tex", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": load_traffic(N)[0],
                         "traffic_unit": "bytes/launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)", "traffic_source": load_traffic(N)[1],
                         "algorithmic_bytes_per_launch": algo_bytes * N,
                         "kernel": "k1_ncc<7,4>", "algorithmic_bytes_per_eval": algo_bytes, "valid_views_per_eval": valid_views,
                         "layout_bytes_per_eval": layout_bytes, "achieved_layout": per_gpu * layout_bytes / 1e9, "frac_layout": per_gpu * layout_bytes / 1e9 / peak,
                         "note": "the binding unit is the L1TEX data pipe (85 % of its peak, issue slots 78 %: profiles/r01_k1_ncc_v4_streamed.txt), not HBM: the config-2 pyramid (154 MB) stays L2-resident and measured DRAM traffic is 1.5 % of the algorithmic bytes; achieved / peak / frac keep the SURVEY 8(d) convention (algorithmic bytes against the measured HBM copy peak); k1_other_configs holds the same kernel on the 1-3 GB pyramids of configs 3 and 5 (latency-bound on DRAM there; the kernel prefetches the next view's footprint into L2: profiles/r02_k1_ncc_config3_pyramid.txt, r02_k1_prefetch_ab.txt)", "l1tex_pct_of_peak": 85.0, "issue_slots_pct": 78.0, "peak_source": peak_src,
                         "wall_ms_timed_region": 1e3 * (t_wall1 - t_wall0)},
        }
        if pipelines:
            main_pipe = next((e for e in pipelines if e["config_id"] == args.config), None)
            if main_pipe is not None:
                out["pipeline"] = main_pipe
                if world == 1 and "error" not in main_pipe and not args.no_cpu_baseline:
                    try:
                        main_pipe["cpu_baseline"] = cpu_reference_sweep_calls_per_sec(scene, args.config, args.scale)
                    except Exception as exc:
                        main_pipe["cpu_baseline"] = {"error": str(exc)}
            out["pipelines"] = pipelines
        if k1_other:
            out["k1_other_configs"] = k1_other
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
