"""Seeded synthetic multi-view scenes in the reference's on-disk layout.

The reference (imkaywu/MVSKit) ships no data; BASELINE.json's configs are "shaped like" public
datasets.  This module renders textured analytic surfaces from calibrated pinhole cameras and
writes exactly what the reference driver reads:

    <prefix>option                  key/value file parsed by Option::init        (pmmvps/option.cpp:35-149)
    <prefix>txt/%08d.txt            "CONTOUR" + 12 floats, row-major 3x4 P       (image/camera.cpp:27-63,109-114)
    <prefix>image/%04d0000.jpg      binary PPM (P6) payload under the .jpg name  (image/photoSet.cpp:32-49, image/image.cpp:827-831)
    <prefix>ply/00000000.patch      PMVS-style seed patches                      (pmmvps/patch.cpp:31-56, patch_manager.cpp:435-466)

The same directory feeds the product's C++ host (mvskit_b200/host) and the oracle build of the
reference (oracle/_ref/libpmref.so, whose CImg stand-in decodes the PPM payload), so both see
bit-identical u8 pixels and float32 projection matrices.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

# ---------------------------------------------------------------------------------------------------
# texture atlas
# ---------------------------------------------------------------------------------------------------

_ATLAS_CACHE: Dict[Tuple[int, int], np.ndarray] = {}


def value_noise_atlas(size: int = 2048, seed: int = 1234, octaves: int = 6, base: int = 8) -> np.ndarray:
    """3-channel periodic value noise, `octaves` octaves, amplitude 1/f, mapped to u8 [16, 240]."""
    key = (size, seed)
    if key in _ATLAS_CACHE:
        return _ATLAS_CACHE[key]
    rng = np.random.RandomState(seed)
    acc = np.zeros((size, size, 3), np.float32)
    coords = np.arange(size, dtype=np.float32)
    for k in range(octaves):
        g = base << k
        grid = rng.rand(g, g, 3).astype(np.float32)
        t = coords * (g / size)
        i0 = np.floor(t).astype(np.int64) % g
        i1 = (i0 + 1) % g
        f = t - np.floor(t)
        f = f * f * (3.0 - 2.0 * f)  # smoothstep
        rows = grid[i0] * (1.0 - f)[:, None, None] + grid[i1] * f[:, None, None]        # (size, g, 3)
        full = rows[:, i0] * (1.0 - f)[None, :, None] + rows[:, i1] * f[None, :, None]   # (size, size, 3)
        acc += full * (0.5 ** k)
    lo, hi = acc.min(), acc.max()
    atlas = (16.0 + (acc - lo) / (hi - lo) * 224.0).astype(np.float32)
    _ATLAS_CACHE[key] = atlas
    return atlas


def _sample_atlas(atlas: np.ndarray, u: np.ndarray, v: np.ndarray) -> np.ndarray:
    """Bilinear, wrapping lookup; u, v in atlas pixels."""
    n = atlas.shape[0]
    u0 = np.floor(u)
    v0 = np.floor(v)
    fu = (u - u0).astype(np.float32)[..., None]
    fv = (v - v0).astype(np.float32)[..., None]
    iu0 = u0.astype(np.int64) % n
    iv0 = v0.astype(np.int64) % n
    iu1 = (iu0 + 1) % n
    iv1 = (iv0 + 1) % n
    top = atlas[iv0, iu0] * (1.0 - fu) + atlas[iv0, iu1] * fu
    bot = atlas[iv1, iu0] * (1.0 - fu) + atlas[iv1, iu1] * fu
    return top * (1.0 - fv) + bot * fv


# ---------------------------------------------------------------------------------------------------
# surfaces
# ---------------------------------------------------------------------------------------------------


class HeightField:
    """z = h(x, y); textured by (x, y).  h must be smooth and gentle (|grad| < ~0.5)."""

    def __init__(self, h: Callable, grad: Callable, extent: float, zrange: float):
        self.h, self.grad, self.extent, self.zrange = h, grad, extent, zrange

    def intersect(self, o: np.ndarray, d: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """o (…,3) origins, d (…,3) unit directions -> (t, hit mask)."""
        dz = d[..., 2]
        safe = np.where(np.abs(dz) < 1e-9, -1e-9, dz)
        t = -o[..., 2] / safe
        for _ in range(12):
            p = o + t[..., None] * d
            g = self.h(p[..., 0], p[..., 1]) - p[..., 2]
            gx, gy = self.grad(p[..., 0], p[..., 1])
            dg = gx * d[..., 0] + gy * d[..., 1] - dz
            dg = np.where(np.abs(dg) < 1e-9, -1e-9, dg)
            t = t - g / dg
        return t, t > 0

    def normal(self, p: np.ndarray) -> np.ndarray:
        gx, gy = self.grad(p[..., 0], p[..., 1])
        n = np.stack([-gx, -gy, np.ones_like(gx)], -1)
        return n / np.linalg.norm(n, axis=-1, keepdims=True)

    def uv(self, p: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        return p[..., 0], p[..., 1]

    def scene_scale(self) -> float:
        return math.sqrt(2 * (2 * self.extent) ** 2 + self.zrange ** 2)


class SphereOnFloor:
    """Unit-ish sphere centred at the origin resting above a textured floor z = -r (occlusions)."""

    def __init__(self, r: float = 1.0):
        self.r = r

    def intersect(self, o, d):
        b = np.sum(o * d, -1)
        c = np.sum(o * o, -1) - self.r ** 2
        disc = b * b - c
        hit_s = disc > 0
        ts = -b - np.sqrt(np.where(hit_s, disc, 0.0))
        hit_s &= ts > 0
        dz = np.where(np.abs(d[..., 2]) < 1e-9, -1e-9, d[..., 2])
        tf = (-self.r - o[..., 2]) / dz
        hit_f = tf > 0
        t = np.where(hit_s, ts, np.where(hit_f, tf, -1.0))
        return t, hit_s | hit_f

    def _on_sphere(self, p):
        return np.abs(np.linalg.norm(p, axis=-1) - self.r) < 1e-3 * self.r

    def normal(self, p):
        s = self._on_sphere(p)
        n = p / np.maximum(np.linalg.norm(p, axis=-1, keepdims=True), 1e-12)
        up = np.zeros_like(p)
        up[..., 2] = 1.0
        return np.where(s[..., None], n, up)

    def uv(self, p):
        s = self._on_sphere(p)
        lon = np.arctan2(p[..., 1], p[..., 0]) * self.r * 2.0
        lat = np.arcsin(np.clip(p[..., 2] / self.r, -1, 1)) * self.r * 2.0
        return np.where(s, lon, p[..., 0] + 7.3), np.where(s, lat, p[..., 1] + 3.1)

    def scene_scale(self) -> float:
        return 2.0 * self.r * math.sqrt(3.0)


def _bumps(amp: float, freq: float):
    def h(x, y):
        return amp * (np.sin(freq * x) * np.cos(0.8 * freq * y) + 0.5 * np.sin(1.7 * freq * x + 0.9 * freq * y))

    def grad(x, y):
        gx = amp * (freq * np.cos(freq * x) * np.cos(0.8 * freq * y) + 0.5 * 1.7 * freq * np.cos(1.7 * freq * x + 0.9 * freq * y))
        gy = amp * (-0.8 * freq * np.sin(freq * x) * np.sin(0.8 * freq * y) + 0.5 * 0.9 * freq * np.cos(1.7 * freq * x + 0.9 * freq * y))
        return gx, gy

    return h, grad


def _plane():
    return (lambda x, y: np.zeros_like(x)), (lambda x, y: (np.zeros_like(x), np.zeros_like(x)))


# ---------------------------------------------------------------------------------------------------
# cameras
# ---------------------------------------------------------------------------------------------------


def look_at_P(eye: np.ndarray, target: np.ndarray, f: float, cx: float, cy: float, up=(0.0, 1.0, 0.0)) -> np.ndarray:
    """3x4 float32 projection of a pinhole camera at `eye` looking at `target`, image y down."""
    z = target - eye
    z = z / np.linalg.norm(z)
    up = np.asarray(up, np.float64)
    x = np.cross(z, up)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z])
    t = -R @ eye
    K = np.array([[f, 0, cx], [0, f, cy], [0, 0, 1.0]])
    return (K @ np.concatenate([R, t[:, None]], 1)).astype(np.float32)


def _arc(n, radius, span_deg):
    """n cameras on an arc about the y axis, slightly off-axis in y so no view is degenerate."""
    out = []
    for a in np.linspace(-span_deg, span_deg, n):
        az = math.radians(a)
        e = np.array([math.sin(az), -0.25, math.cos(az)])
        out.append(radius * e / np.linalg.norm(e))
    return out


def _ring(n, radius, elev_deg):
    el = math.radians(elev_deg)
    return [np.array([radius * math.cos(el) * math.cos(2 * math.pi * k / n), radius * math.cos(el) * math.sin(2 * math.pi * k / n), radius * math.sin(el)]) for k in range(n)]


def _dome(nx, ny, radius, span_deg):
    out = []
    for j in np.linspace(-span_deg, span_deg, ny):
        for i in np.linspace(-span_deg, span_deg, nx):
            ax, ay = math.radians(i), math.radians(j)
            d = np.array([math.tan(ax), math.tan(ay), 1.0])
            out.append(radius * d / np.linalg.norm(d))
    return out


# ---------------------------------------------------------------------------------------------------
# scene
# ---------------------------------------------------------------------------------------------------


@dataclass
class Scene:
    name: str
    width: int
    height: int
    f: float
    P: np.ndarray                      # (n, 3, 4) float32, level 0
    eyes: np.ndarray                   # (n, 3) float64
    surface: object
    tex_scale: float                   # atlas pixels per scene unit
    level: int = 1
    csize: int = 2
    wsize: int = 7
    min_image_num: int = 3
    threshold: float = 0.7
    iters: int = 3
    images: List[np.ndarray] = field(default_factory=list)   # (H, W, 3) uint8, filled by render()
    masks: List[Optional[np.ndarray]] = field(default_factory=list)   # (H, W) uint8 grey or None per view, filled by make_masks()
    mask_formats: List[str] = field(default_factory=list)             # "pgm" | "pbm" per view (how write_scene stores them)

    @property
    def nviews(self) -> int:
        return int(self.P.shape[0])

    @property
    def scene_scale(self) -> float:
        return float(self.surface.scene_scale())

    # -- geometry helpers (float64; data generation only) ------------------------------------------
    def _KR(self, v):
        M = self.P[v, :, :3].astype(np.float64)
        return M

    def rays(self, v: int, px: np.ndarray, py: np.ndarray) -> np.ndarray:
        """Unit ray directions through level-0 pixels (px, py) of view v."""
        Minv = np.linalg.inv(self._KR(v))
        d = np.stack([px, py, np.ones_like(px)], -1) @ Minv.T
        return d / np.linalg.norm(d, axis=-1, keepdims=True)

    def project(self, v: int, X: np.ndarray) -> np.ndarray:
        Xh = np.concatenate([X, np.ones(X.shape[:-1] + (1,))], -1)
        q = Xh @ self.P[v].astype(np.float64).T
        return np.concatenate([q[..., :2] / q[..., 2:3], q[..., 2:3]], -1)

    def cast(self, v: int, px: np.ndarray, py: np.ndarray):
        o = np.broadcast_to(self.eyes[v], px.shape + (3,))
        d = self.rays(v, px, py)
        t, hit = self.surface.intersect(o, d)
        return o + t[..., None] * d, hit

    def visible(self, v: int, X: np.ndarray, tol: float = 1e-3) -> np.ndarray:
        """True where X is the first surface hit along the ray from camera v (and inside the image)."""
        q = self.project(v, X)
        inside = (q[..., 2] > 0) & (q[..., 0] >= 8) & (q[..., 0] < self.width - 9) & (q[..., 1] >= 8) & (q[..., 1] < self.height - 9)
        d = X - self.eyes[v]
        dist = np.linalg.norm(d, axis=-1)
        d = d / np.maximum(dist[..., None], 1e-12)
        t, hit = self.surface.intersect(np.broadcast_to(self.eyes[v], X.shape), d)
        return inside & hit & (np.abs(t - dist) < tol * self.scene_scale)

    def render(self, views: Optional[List[int]] = None) -> "Scene":
        atlas = value_noise_atlas()
        self.images = [None] * self.nviews if not self.images else self.images
        ys, xs = np.mgrid[0:self.height, 0:self.width].astype(np.float64)
        for v in (range(self.nviews) if views is None else views):
            X, hit = self.cast(v, xs, ys)
            u, w = self.surface.uv(X)
            col = _sample_atlas(atlas, u * self.tex_scale, w * self.tex_scale)
            sky = 96.0 + 24.0 * np.sin(xs * 0.05 + v)[..., None] * np.cos(ys * 0.043)[..., None] * np.ones(3)
            img = np.where(hit[..., None], col, sky)
            self.images[v] = np.clip(np.floor(img + 0.5), 0, 255).astype(np.uint8)
        return self

    def make_masks(self, seed: int = 77, maskless=(2,), pbm=(1,)) -> "Scene":
        """Silhouette masks in the reference's sense (image.cpp:143-161: grey > 127 = inside): one ellipse per view with a
        ragged grey rim (values on both sides of the 127 threshold, odd positions so 2x2 pyramid blocks straddle the edge).
        Views in `maskless` get none (Image::getMask = -1 there), views in `pbm` are stored as binary P4."""
        rng = np.random.default_rng(seed)
        ys, xs = np.mgrid[0:self.height, 0:self.width].astype(np.float64)
        self.masks, self.mask_formats = [], []
        for v in range(self.nviews):
            if v in maskless:
                self.masks.append(None); self.mask_formats.append("pgm")
                continue
            cx = self.width * (0.5 + 0.06 * math.cos(1.3 * v)); cy = self.height * (0.5 + 0.05 * math.sin(0.9 * v))
            rx = self.width * (0.31 + 0.015 * v); ry = self.height * (0.33 + 0.01 * v)
            d = np.sqrt(((xs - cx) / rx) ** 2 + ((ys - cy) / ry) ** 2)
            grey = np.clip(255.0 * (1.06 - d) / 0.12, 0, 255)                 # ramp 255 -> 0 across the rim
            grey = np.clip(grey + rng.integers(-40, 41, size=grey.shape), 0, 255)
            m = np.floor(grey + 0.5).astype(np.uint8)
            if v in pbm:
                m = np.where(m > 127, 255, 0).astype(np.uint8)
                self.mask_formats.append("pbm")
            else:
                self.mask_formats.append("pgm")
            self.masks.append(m)
        return self

    # -- hypotheses for the NCC micro-benchmark / parity tests -----------------------------------------
    def neighbours(self, X: np.ndarray, n: np.ndarray, ref: int, k: int, cos_min: float = math.cos(math.radians(58.0))) -> np.ndarray:
        """For each point: [ref] + the k-1 other views with the smallest angle to the ref ray that see it."""
        N = X.shape[0]
        ray_ref = self.eyes[ref] - X
        ray_ref /= np.linalg.norm(ray_ref, axis=-1, keepdims=True)
        score = np.full((N, self.nviews), np.inf)
        for v in range(self.nviews):
            if v == ref:
                continue
            r = self.eyes[v] - X
            r /= np.linalg.norm(r, axis=-1, keepdims=True)
            ok = self.visible(v, X) & (np.sum(r * n, -1) > cos_min)
            score[:, v] = np.where(ok, 1.0 - np.sum(r * ray_ref, -1), np.inf)
        m = min(k - 1, self.nviews - 1)
        order = np.argsort(score, axis=1, kind="stable")[:, :m]
        good = np.take_along_axis(score, order, 1) < np.inf
        out = np.full((N, k), -1, np.int32)
        out[:, 0] = ref
        for j in range(m):
            out[:, j + 1] = np.where(good[:, j], order[:, j], -1)
        return out

    def hypotheses(self, n: int, seed: int = 7, depth_jitter: float = 0.02, normal_jitter_deg: float = 20.0, tau: int = 6,
                   order: str = "random", well_observed: bool = True):
        """n hypotheses around ground truth, each with up to tau views ([0] = reference).

        order="random": random pixels of each reference view in draw order; order="grid": the same draws sorted
        in Z-order of their reference pixel, i.e. the spatially coherent order in which a patch grid is walked.
        well_observed=True keeps only pixels whose ground-truth point faces the reference view and is seen by
        tau-1 other views (every eval then samples all tau views); False keeps every draw, so the set also holds
        hypotheses with missing, back-facing or out-of-image views (edge cases for the parity tests).
        Returns coord (n,4) f32, normal (n,4) f32, views (n,tau) i32 (-1 padded, compacted), nviews (n,) i32.
        """
        rng = np.random.RandomState(seed)
        per = (n + self.nviews - 1) // self.nviews
        coords, normals, views = [], [], []
        cos_ref, cos_nb = math.cos(math.radians(35.0)), math.cos(math.radians(42.0))
        for v in range(self.nviews):
            m = min(per, n - v * per)
            if m <= 0:
                break
            # rejection-sample pixels whose ground-truth point is well observed: facing the reference view
            # and seen by tau-1 other views (so the jittered hypothesis normally keeps all tau views)
            got_px, got_py, have, rounds = [], [], 0, 0
            while have < m and rounds < 12:
                k = max(1024, 3 * (m - have))
                px = rng.uniform(24, self.width - 25, k)
                py = rng.uniform(24, self.height - 25, k)
                X, hit = self.cast(v, px, py)
                nrm = self.surface.normal(X)
                ray = self.eyes[v] - X
                ray /= np.linalg.norm(ray, axis=-1, keepdims=True)
                good = hit & (np.sum(ray * nrm, -1) > cos_ref)
                if not well_observed:
                    good = np.ones_like(hit)
                elif rounds < 11:
                    idx = np.nonzero(good)[0]
                    if idx.size:
                        vw = self.neighbours(X[idx], nrm[idx], v, tau, cos_min=cos_nb)
                        good[idx] &= (vw >= 0).sum(1) >= min(tau, self.nviews)
                else:
                    good = hit                       # give up filtering: keep the batch size exact
                got_px.append(px[good]); got_py.append(py[good])
                have += int(good.sum()); rounds += 1
            px, py = np.concatenate(got_px)[:m], np.concatenate(got_py)[:m]
            if px.size < m:                          # pathological scene: pad with unfiltered draws
                px = np.concatenate([px, rng.uniform(24, self.width - 25, m - px.size)])
                py = np.concatenate([py, rng.uniform(24, self.height - 25, m - py.size)])
            if order == "grid":
                code = np.zeros(m, np.int64)
                ix, iy = px.astype(np.int64), py.astype(np.int64)
                for b in range(13):
                    code |= ((ix >> b) & 1) << (2 * b) | ((iy >> b) & 1) << (2 * b + 1)
                o = np.argsort(code, kind="stable")
                px, py = px[o], py[o]
            X, hit = self.cast(v, px, py)
            nrm = self.surface.normal(X)
            vw = self.neighbours(X, nrm, v, tau, cos_min=cos_nb)
            d = X - self.eyes[v]
            X = self.eyes[v] + d * (1.0 + rng.uniform(-depth_jitter, depth_jitter, (m, 1)))
            ang = math.radians(normal_jitter_deg)
            pert = rng.uniform(-1, 1, (m, 3)) * math.tan(ang) * 0.6
            nj = nrm + pert - np.sum(pert * nrm, -1, keepdims=True) * nrm
            nj /= np.linalg.norm(nj, axis=-1, keepdims=True)
            X[~hit] = 0.0
            coords.append(X), normals.append(nj), views.append(vw)
        X = np.concatenate(coords)[:n]
        nj = np.concatenate(normals)[:n]
        vw = np.concatenate(views)[:n]
        coord = np.concatenate([X, np.ones((X.shape[0], 1))], 1).astype(np.float32)
        normal = np.concatenate([nj, np.zeros((X.shape[0], 1))], 1).astype(np.float32)
        # compact the -1 holes to the tail
        key = np.where(vw >= 0, 0, 1)
        order = np.argsort(key, axis=1, kind="stable")
        vw = np.take_along_axis(vw, order, 1).astype(np.int32)
        nviews = (vw >= 0).sum(1).astype(np.int32)
        return coord, normal, vw, nviews

    # -- seed patches (ply/00000000.patch) ----------------------------------------------------------
    def seeds(self, stride: int = 4, seed: int = 42, depth_noise: float = 0.005, normal_noise_deg: float = 10.0, max_images: int = 12,
              views: Optional[List[int]] = None):
        """Seed patches around ground truth, one per `stride`-th cell of every view.  views = None: one random stream over all views
        in order (configs 1 and 2 as used since round 1).  views = [...]: only those views, each with its OWN stream (seed, view) --
        what the parallel generator for the big configs uses, so the result does not depend on how the views are split up."""
        rng = np.random.RandomState(seed)
        sc = 1 << self.level
        gw = (self.width // sc + self.csize - 1) // self.csize
        gh = (self.height // sc + self.csize - 1) // self.csize
        recs = []
        for v in (range(self.nviews) if views is None else views):
            if views is not None:
                rng = np.random.RandomState((seed * 7919 + 104729 * (v + 1)) % (2 ** 31))
            cx, cy = np.meshgrid(np.arange(6, gw - 6, stride), np.arange(6, gh - 6, stride))
            cx, cy = cx.ravel().astype(np.float64), cy.ravel().astype(np.float64)
            u1 = (self.csize * (2 * cx + 1) - 1) / 2.0
            v1 = (self.csize * (2 * cy + 1) - 1) / 2.0
            X, hit = self.cast(v, u1 * sc, v1 * sc)
            nrm = self.surface.normal(X)
            ray = self.eyes[v] - X
            dist = np.linalg.norm(ray, axis=-1, keepdims=True)
            ray /= dist
            facing = np.sum(ray * nrm, -1) > math.cos(math.radians(50.0))
            vw = self.neighbours(X, nrm, v, max_images)
            keep = hit & facing & ((vw >= 0).sum(1) >= self.min_image_num)
            X = X - ray * rng.normal(0.0, depth_noise * self.scene_scale, (X.shape[0], 1))
            pert = rng.normal(0, math.tan(math.radians(normal_noise_deg)) / math.sqrt(2.0), X.shape)
            nj = nrm + pert - np.sum(pert * nrm, -1, keepdims=True) * nrm
            nj /= np.linalg.norm(nj, axis=-1, keepdims=True)
            unit = 2.0 * dist[:, 0] * sc / (2.0 * self.f)
            for i in np.nonzero(keep)[0]:
                ids = [int(a) for a in vw[i] if a >= 0]
                recs.append((X[i], nj[i], float(2.0 * unit[i]), ids))
        return recs


def _render_job(args):
    config, scale, nviews, views = args
    sc = make_scene(config, scale=scale, nviews=nviews)
    sc.render(views=views)
    return [(v, sc.images[v]) for v in views]


def _seed_job(args):
    config, scale, nviews, views, kw = args
    return make_scene(config, scale=scale, nviews=nviews).seeds(views=views, **kw)


def _pool_map(fn, jobs, procs):
    if procs <= 1 or len(jobs) <= 1:
        return [fn(j) for j in jobs]
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(min(procs, len(jobs))) as pool:
        return pool.map(fn, jobs)


def render_views(config: int, scale: float, views: List[int], procs: int, nviews: Optional[int] = None):
    """[(view, image)] for `views`, rendered by `procs` processes (every view renders independently and deterministically)."""
    procs = max(1, min(procs, len(views)))
    jobs = [(config, scale, nviews, views[i::procs]) for i in range(procs)]
    return [vi for part in _pool_map(_render_job, jobs, procs) for vi in part]


def seed_records(config: int, scale: float, views: List[int], procs: int, nviews: Optional[int] = None, **kw):
    """Scene.seeds(views=...) over `procs` processes; records come back grouped by view in ascending view order."""
    procs = max(1, min(procs, len(views)))
    chunks = [views[i::procs] for i in range(procs)]
    parts = _pool_map(_seed_job, [(config, scale, nviews, ch, kw) for ch in chunks], procs)
    # each chunk's records are in its own view order; regroup into ascending view order
    by_view = {}
    for ch, recs in zip(chunks, parts):
        ref_of = [r[3][0] for r in recs]
        for v in ch:
            by_view[v] = [r for r, rv in zip(recs, ref_of) if rv == v]
    return [r for v in sorted(by_view) for r in by_view[v]]


def seed_arrays(scene: "Scene", recs=None, **kw):
    """scene.seeds() as the record arrays pmk_store_add takes: coord4, normal4, scal4 = (ncc 1, dscale, 0, 0), images, nimages."""
    if recs is None:
        recs = scene.seeds(**kw)
    n, V = len(recs), scene.nviews
    coord, normal, scal = np.ones((n, 4), np.float32), np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32)
    images, nimg = np.zeros((n, V), np.int32), np.zeros(n, np.int32)
    for i, (X, N, ds, ids) in enumerate(recs):
        coord[i, :3], normal[i, :3] = X, N
        scal[i] = (1.0, ds, 0.0, 0.0)
        images[i, :len(ids)] = ids
        nimg[i] = len(ids)
    return coord, normal, scal, images, nimg


def make_scene(config: int, scale: float = 1.0, nviews: Optional[int] = None) -> Scene:
    """BASELINE.json configs 1..5.  `scale` shrinks the image size (tests)."""
    def dims(w, h, f):
        return max(64, int(w * scale)) // 4 * 4, max(48, int(h * scale)) // 4 * 4, f * scale

    if config == 1:      # textured plane, 5 views 640x480, test.cpp:96-100 intrinsics
        W, H, f = dims(640, 480, 765.70)
        eyes = _arc(nviews or 5, 3.0, 20.0)
        h, g = _plane()
        surf = HeightField(h, g, extent=1.0, zrange=0.0)
        target, iters = np.zeros(3), 3
    elif config == 2:    # templeRing-shaped: sphere on a floor, 47 views on a ring
        W, H, f = dims(640, 480, 765.70)
        eyes = _ring(nviews or 47, 4.0, 30.0)
        surf = SphereOnFloor(1.0)
        target, iters = np.zeros(3), 3
    elif config == 3:    # DTU-shaped: undulating height field, 7x7 dome, 1600x1200
        W, H, f = dims(1600, 1200, 2890.0)
        n = nviews or 49
        side = int(round(math.sqrt(n)))
        eyes = _dome(side, (n + side - 1) // side, 3.0, 35.0)[:n]
        h, g = _bumps(0.08, 3.0)
        surf = HeightField(h, g, extent=0.8, zrange=0.25)
        target, iters = np.zeros(3), 4
    elif config == 4:    # fountain-P11-shaped: relief, 11 views on an arc, 3072x2048
        W, H, f = dims(3072, 2048, 2760.0)
        eyes = _arc(nviews or 11, 3.0, 30.0)
        h, g = _bumps(0.12, 2.2)
        surf = HeightField(h, g, extent=1.6, zrange=0.4)
        target, iters = np.zeros(3), 3
    elif config == 5:    # large scene: 16x8 dome, 1920x1080
        W, H, f = dims(1920, 1080, 1600.0)
        n = nviews or 128
        eyes = _dome(16, 8, 3.0, 40.0)[:n]
        h, g = _bumps(0.10, 2.5)
        surf = HeightField(h, g, extent=1.8, zrange=0.3)
        target, iters = np.zeros(3), 3
    else:
        raise ValueError("config must be 1..5")
    eyes = np.asarray(eyes, np.float64)
    up = (0.0, 0.0, 1.0) if config == 2 else (0.0, 1.0, 0.0)
    P = np.stack([look_at_P(e, target, f, W / 2.0, H / 2.0, up) for e in eyes])
    # finest atlas octave (8 atlas px) ~ 2 level-0 pixels at the mean viewing distance
    dist = float(np.mean(np.linalg.norm(eyes - target, axis=1)))
    tex_scale = 8.0 / (2.0 * dist / f)
    return Scene(name=f"config{config}", width=W, height=H, f=f, P=P, eyes=eyes, surface=surf, tex_scale=tex_scale, iters=iters)


# ---------------------------------------------------------------------------------------------------
# on-disk layout
# ---------------------------------------------------------------------------------------------------


def _f32(x: float) -> str:
    return "%.9g" % float(np.float32(x))


def write_scene(scene: Scene, prefix: str, with_seeds: bool = True, seed_stride: int = 4) -> str:
    """Write the reference layout under `prefix` (created; a trailing '/' is appended).  Returns prefix."""
    if not prefix.endswith("/"):
        prefix += "/"
    for d in ("txt", "image", "ply"):
        os.makedirs(prefix + d, exist_ok=True)
    if not scene.images or any(im is None for im in scene.images):
        scene.render()
    n = scene.nviews
    with open(prefix + "option", "w") as fh:
        fh.write(f"# synthetic {scene.name}\nimage {n}\nillum 1\nlevel {scene.level}\ncsize {scene.csize}\n"
                 f"wsize {scene.wsize}\nthreshold {scene.threshold}\nminImageNum {scene.min_image_num}\nimages -1 0 {n}\n")
    for v in range(n):
        with open(prefix + "txt/%08d.txt" % v, "w") as fh:
            fh.write("CONTOUR\n")
            for r in range(3):
                fh.write(" ".join(_f32(x) for x in scene.P[v, r]) + "\n")
        im = scene.images[v]
        with open(prefix + "image/%04d0000.jpg" % v, "wb") as fh:
            fh.write(b"P6\n%d %d\n255\n" % (im.shape[1], im.shape[0]))
            fh.write(np.ascontiguousarray(im).tobytes())
    if scene.masks:
        # <prefix>mask/%08d.pgm (binary P5) or .pbm (binary P4; the reference reads the bits as ONE stream without the
        # per-row padding of the format, bit set = outside, image.cpp:881-943) -- photoSet.cpp:48, image.cpp:82-89,143-161
        os.makedirs(prefix + "mask", exist_ok=True)
        for v, m in enumerate(scene.masks):
            if m is None:
                continue
            if scene.mask_formats[v] == "pbm":
                bits = np.packbits((np.ascontiguousarray(m).reshape(-1) <= 127).astype(np.uint8))
                with open(prefix + "mask/%08d.pbm" % v, "wb") as fh:
                    fh.write(b"P4\n%d %d\n" % (m.shape[1], m.shape[0]))
                    fh.write(bits.tobytes() + b"\0")
            else:
                with open(prefix + "mask/%08d.pgm" % v, "wb") as fh:
                    fh.write(b"P5\n%d %d\n255\n" % (m.shape[1], m.shape[0]))
                    fh.write(np.ascontiguousarray(m).tobytes())
    if with_seeds:
        recs = scene.seeds(stride=seed_stride)
        with open(prefix + "ply/00000000.patch", "w") as fh:
            fh.write("PATCHES\n%d\n" % len(recs))
            for X, nrm, dscale, ids in recs:
                fh.write("PATCHES\n%s %s %s 1\n%s %s %s 0\n-1 %s 0.1\n%d\n%s\n0\n\n\n" % (
                    _f32(X[0]), _f32(X[1]), _f32(X[2]), _f32(nrm[0]), _f32(nrm[1]), _f32(nrm[2]),
                    _f32(dscale), len(ids), " ".join(str(i) for i in ids)))
    return prefix
