"""ctypes binding of the C ABI in include/pmk.h (mvskit_b200/libpmk.so).

This is plumbing for tests and bench.py; the product is the shared library.  There is no CPU
fallback: importing works anywhere (so the symbol table can be checked on a CPU box), but creating
a :class:`Context` raises unless a CUDA device is present, and a missing library raises at load.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpmk.so")

PMK_MAX_LEVELS = 6
PMK_MAX_TAU = 8


class PmkError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("device", C.c_int), ("nviews", C.c_int), ("level", C.c_int), ("csize", C.c_int), ("wsize", C.c_int),
                ("min_image_num", C.c_int), ("ncc_threshold", C.c_float), ("max_angle_threshold", C.c_float),
                ("quad_threshold", C.c_float), ("max_patches", C.c_int), ("cell_capacity", C.c_int), ("jitter_mode", C.c_int), ("sweep_group", C.c_int)]


class Thresholds(C.Structure):
    _fields_ = [("tau", C.c_int), ("depth", C.c_int)] + [(k, C.c_float) for k in (
        "ncc_threshold", "ncc_threshold_before", "angle_threshold0", "angle_threshold1", "max_angle_threshold",
        "quad_threshold", "neighbor_threshold", "neighbor_threshold1", "neighbor_threshold2")]


class Camera(C.Structure):
    _fields_ = [("P", C.c_float * 12), ("center", C.c_float * 4), ("oaxis", C.c_float * 4), ("xaxis", C.c_float * 3),
                ("yaxis", C.c_float * 3), ("zaxis", C.c_float * 3), ("ipscale", C.c_float)]


class ForcedIO(C.Structure):
    """pmk_forced_io (include/pmk.h)."""
    _fields_ = [("ntries", C.c_int), ("stride", C.c_int)] + [(n, C.c_void_p) for n in (
        "code", "ncc0", "coord4", "normal4", "scal4", "nimages", "images", "ntries_out", "outcome", "branch_full", "post_ret",
        "nimages_out", "images_out", "grids_out", "nvimages_out", "vimages_out", "vgrids_out", "tmp_out")]


_lib = None


def lib() -> C.CDLL:
    """Load libpmk.so.  Raises (never falls back) when the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PmkError(f"{LIB_PATH} is missing: build it with `python -m mvskit_b200.build` (no CPU fallback exists)")
        _lib = C.CDLL(LIB_PATH)
        _lib.pmk_last_error.restype = C.c_char_p
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _chk(rc: int):
    if rc != 0:
        raise PmkError(f"pmk error {rc}: {lib().pmk_last_error().decode()}")


class DeviceBuffer:
    def __init__(self, ctx: "Context", nbytes: int):
        self.ctx, self.nbytes = ctx, int(nbytes)
        self.ptr = C.c_void_p()
        _chk(lib().pmk_device_alloc(ctx.h, C.c_uint64(self.nbytes), C.byref(self.ptr)))

    def upload(self, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        _chk(lib().pmk_memcpy_h2d(self.ctx.h, self.ptr, _p(arr), C.c_uint64(arr.nbytes)))
        return self

    def download(self, arr: np.ndarray):
        assert arr.flags["C_CONTIGUOUS"] and arr.nbytes <= self.nbytes
        _chk(lib().pmk_memcpy_d2h(self.ctx.h, _p(arr), self.ptr, C.c_uint64(arr.nbytes)))
        return arr

    def free(self):
        if self.ptr:
            lib().pmk_device_free(self.ctx.h, self.ptr)
            self.ptr = C.c_void_p()


class Patches:
    """numpy-side patch records in the layout of include/pmk.h (scal = ncc, dscale, ascale, tmp)."""

    def __init__(self, n: int, maxv: int):
        self.n, self.maxv = n, maxv
        self.coord = np.zeros((n, 4), np.float32)
        self.normal = np.zeros((n, 4), np.float32)
        self.scal = np.zeros((n, 4), np.float32)
        self.images = np.full((n, maxv), -1, np.int32)
        self.nimages = np.zeros(n, np.int32)
        self.grids = np.zeros((n, maxv, 2), np.int32)
        self.vimages = np.full((n, maxv), -1, np.int32)
        self.nvimages = np.zeros(n, np.int32)
        self.vgrids = np.zeros((n, maxv, 2), np.int32)


class Context:
    """One per GPU.  Mirrors PmMvps::init's effect on the device (pmmvps/pmmvps.cpp:18-68)."""

    def __init__(self, nviews: int, device: int = 0, level: int = 1, csize: int = 2, wsize: int = 7, min_image_num: int = 3,
                 ncc_threshold: float = 0.7, max_patches: int = 0, cell_capacity: int = 0, jitter_mode: int = 0, sweep_group: int = 1):
        L = lib()
        cfg = Config()
        L.pmk_default_config(C.byref(cfg))
        cfg.device, cfg.nviews, cfg.level, cfg.csize, cfg.wsize = device, nviews, level, csize, wsize
        cfg.min_image_num, cfg.ncc_threshold, cfg.max_patches = min_image_num, ncc_threshold, max_patches
        cfg.cell_capacity, cfg.jitter_mode, cfg.sweep_group = cell_capacity, jitter_mode, sweep_group
        self.cfg = cfg
        self.h = C.c_void_p()
        _chk(L.pmk_create(C.byref(cfg), C.byref(self.h)))
        self.nviews, self.level, self.nlevels = nviews, level, level + 3
        self.tau = min(2 * min_image_num, nviews)

    def close(self):
        if self.h:
            lib().pmk_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- scene ---------------------------------------------------------------------------------------
    def set_view(self, view: int, P: np.ndarray, rgb: np.ndarray):
        P = np.ascontiguousarray(P, np.float32).reshape(12)
        rgb = np.ascontiguousarray(rgb, np.uint8)
        assert rgb.ndim == 3 and rgb.shape[2] == 3
        _chk(lib().pmk_set_view(self.h, view, _p(P), _p(rgb), rgb.shape[1], rgb.shape[0]))

    def set_view_jpeg(self, view: int, P: np.ndarray, jpeg: bytes):
        P = np.ascontiguousarray(P, np.float32).reshape(12)
        buf = np.frombuffer(jpeg, np.uint8)
        w, h = C.c_int(), C.c_int()
        _chk(lib().pmk_set_view_jpeg(self.h, view, _p(P), _p(buf), C.c_uint64(len(buf)), C.byref(w), C.byref(h)))
        return w.value, h.value

    def set_view_mask(self, view: int, grey: np.ndarray):
        grey = np.ascontiguousarray(grey, np.uint8)
        assert grey.ndim == 2
        _chk(lib().pmk_set_view_mask(self.h, view, _p(grey), grey.shape[1], grey.shape[0]))

    def level_mask(self, view: int, level: int):
        """Image::m_masks[level] (None when the view has no mask)."""
        w, h = self.level_dims(view, level)
        out = np.empty((h, w), np.uint8)
        has = C.c_int()
        _chk(lib().pmk_get_level_mask(self.h, view, level, _p(out), C.byref(has)))
        return out if has.value else None

    def probe_mask(self, coord, view: int = -1) -> np.ndarray:
        """PhotoSet::getMask(coord, m_level) (view < 0) or PhotoSet::getMask(view, coord, m_level)."""
        coord = np.ascontiguousarray(coord, np.float32)
        out = np.empty(len(coord), np.int32)
        _chk(lib().pmk_probe_mask(self.h, len(coord), view, _p(coord), _p(out)))
        return out

    def set_scene(self, P: np.ndarray, images: Sequence[np.ndarray], masks=None):
        for v in range(self.nviews):
            self.set_view(v, P[v], images[v])
            if masks and masks[v] is not None:
                self.set_view_mask(v, masks[v])

    def thresholds(self) -> Thresholds:
        t = Thresholds()
        _chk(lib().pmk_get_thresholds(self.h, C.byref(t)))
        return t

    def set_depth(self, d: int):
        _chk(lib().pmk_set_depth(self.h, d))

    def update_threshold(self):
        _chk(lib().pmk_update_threshold(self.h))

    def camera(self, view: int, level: int = 0) -> dict:
        c = Camera()
        _chk(lib().pmk_get_camera(self.h, view, level, C.byref(c)))
        return dict(P=np.array(c.P, np.float32).reshape(3, 4), center=np.array(c.center, np.float32), oaxis=np.array(c.oaxis, np.float32),
                    xaxis=np.array(c.xaxis, np.float32), yaxis=np.array(c.yaxis, np.float32), zaxis=np.array(c.zaxis, np.float32),
                    ipscale=np.float32(c.ipscale))

    def level_dims(self, view: int, level: int):
        w, h = C.c_int(), C.c_int()
        _chk(lib().pmk_get_level_dims(self.h, view, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def grid_dims(self, view: int):
        w, h = C.c_int(), C.c_int()
        _chk(lib().pmk_get_grid_dims(self.h, view, C.byref(w), C.byref(h)))
        return w.value, h.value

    def level_image(self, view: int, level: int) -> np.ndarray:
        w, h = self.level_dims(view, level)
        out = np.zeros((h, w, 3), np.uint8)
        _chk(lib().pmk_get_level_image(self.h, view, level, _p(out)))
        return out

    # -- K1 ------------------------------------------------------------------------------------------
    def ncc_eval(self, coord, normal, views, nviews, want_levels: bool = False):
        """PatchManager::computeNcc for a batch; host arrays in, host arrays out."""
        coord = np.ascontiguousarray(coord, np.float32)
        normal = np.ascontiguousarray(normal, np.float32)
        views = np.ascontiguousarray(views, np.int32)
        nviews = np.ascontiguousarray(nviews, np.int32)
        n = coord.shape[0]
        incc, ncc = np.empty(n, np.float32), np.empty(n, np.float32)
        levels = np.empty((n, self.tau), np.int32) if want_levels else None
        stride = views.shape[1] if views.ndim == 2 else 1
        _chk(lib().pmk_ncc_eval(self.h, n, _p(coord), _p(normal), _p(views), _p(nviews), stride, _p(incc), _p(ncc), _p(levels)))
        return (incc, ncc, levels) if want_levels else (incc, ncc)

    @staticmethod
    def pack_hypotheses(coord, normal, views, nviews):
        """float4 / float4 / int32 records -> the byte-lean layout of pmk_ncc_eval_packed (w = 1 / w = 0 implied, byte view ids)."""
        views, nviews = np.asarray(views), np.asarray(nviews)
        if nviews.min() < 0 or nviews.max() > min(255, views.shape[1]):
            raise PmkError("pack_hypotheses: view counts must lie in [0, min(255, row length)]")
        v8 = np.where((views < 0) | (views > 254), 255, views).astype(np.uint8)      # 255 = "no such view" (the call needs nviews <= 255)
        return (np.ascontiguousarray(np.asarray(coord, np.float32)[:, :3]), np.ascontiguousarray(np.asarray(normal, np.float32)[:, :3]),
                np.ascontiguousarray(v8), np.ascontiguousarray(nviews, np.uint8))

    def ncc_eval_packed(self, coord3, normal3, views8, nviews8, want_levels: bool = False, out=None):
        """pmk_ncc_eval_packed: host arrays (n,3) f32, (n,3) f32, (n,stride) u8, (n,) u8 -> incc, ncc[, levels]."""
        n, stride = coord3.shape[0], views8.shape[1]
        assert coord3.dtype == np.float32 and normal3.dtype == np.float32 and views8.dtype == np.uint8 and nviews8.dtype == np.uint8
        incc, ncc = out if out is not None else (np.empty(n, np.float32), np.empty(n, np.float32))
        levels = np.empty((n, self.tau), np.int32) if want_levels else None
        _chk(lib().pmk_ncc_eval_packed(self.h, n, _p(coord3), _p(normal3), _p(views8), _p(nviews8), stride, _p(incc), _p(ncc), _p(levels)))
        return (incc, ncc, levels) if want_levels else (incc, ncc)

    def ncc_eval_dev(self, n: int, coord: DeviceBuffer, normal: DeviceBuffer, views: DeviceBuffer, nviews: DeviceBuffer, stride: int,
                     incc: DeviceBuffer, ncc: Optional[DeviceBuffer] = None, levels: Optional[DeviceBuffer] = None):
        _chk(lib().pmk_ncc_eval_dev(self.h, n, coord.ptr, normal.ptr, views.ptr, nviews.ptr, stride, incc.ptr,
                                    ncc.ptr if ncc else None, levels.ptr if levels else None))

    # -- candidate kernels ------------------------------------------------------------------------------
    @staticmethod
    def _cv(coord, normal, views, nviews):
        return (np.ascontiguousarray(coord, np.float32), np.ascontiguousarray(normal, np.float32),
                np.ascontiguousarray(views, np.int32), np.ascontiguousarray(nviews, np.int32))

    def set_inccs(self, coord, normal, views, nviews, robust: int, pairwise: bool = False):
        """Optim::setINCCs (optim.cpp:708-783)."""
        coord, normal, views, nviews = self._cv(coord, normal, views, nviews)
        n, stride = views.shape
        out = np.zeros((n, stride, stride) if pairwise else (n, stride), np.float32)
        _chk(lib().pmk_set_inccs(self.h, n, _p(coord), _p(normal), _p(views), _p(nviews), stride, int(robust), int(pairwise), _p(out)))
        return out

    def pre_process(self, coord, normal, views, nviews, maxv: Optional[int] = None):
        """Optim::preProcess (optim.cpp:137-163) -> ret, images (n, maxv), nimages, dscale, ascale."""
        coord, normal, views, nviews = self._cv(coord, normal, views, nviews)
        n, stride = views.shape
        maxv = maxv or self.nviews
        ret, images, nimg = np.zeros(n, np.int32), np.zeros((n, maxv), np.int32), np.zeros(n, np.int32)
        ds, asc = np.zeros(n, np.float32), np.zeros(n, np.float32)
        _chk(lib().pmk_pre_process(self.h, n, _p(coord), _p(normal), _p(views), _p(nviews), stride, maxv, _p(ret), _p(images), _p(nimg), _p(ds), _p(asc)))
        return ret, images, nimg, ds, asc

    def cost_func(self, coord, normal, dscale, views, nviews, patch_of_item, x3):
        """Optim::cost_func (optim.cpp:401-468) at encoded points."""
        coord, normal, views, nviews = self._cv(coord, normal, views, nviews)
        dscale = np.ascontiguousarray(dscale, np.float32)
        pid = np.ascontiguousarray(patch_of_item, np.int32)
        x3 = np.ascontiguousarray(x3, np.float64)
        out = np.zeros(len(pid), np.float64)
        _chk(lib().pmk_cost_func(self.h, len(coord), _p(coord), _p(normal), _p(dscale), _p(views), _p(nviews), views.shape[1], len(pid), _p(pid), _p(x3), _p(out)))
        return out

    def refine(self, coord, normal, dscale, views, nviews, streams, seed: int, trace: bool = False):
        """Optim::refinePatch (optim.cpp:470-547) with the PMR1 schedule."""
        coord, normal, views, nviews = self._cv(coord, normal, views, nviews)
        coord, normal = coord.copy(), normal.copy()
        dscale = np.ascontiguousarray(dscale, np.float32)
        streams = np.ascontiguousarray(streams, np.uint64)
        n = len(coord)
        ncc = np.zeros(n, np.float32)
        tr = np.zeros((n, 97, 4), np.float64) if trace else None
        _chk(lib().pmk_refine(self.h, n, _p(coord), _p(normal), _p(dscale), _p(views), _p(nviews), views.shape[1], _p(streams), C.c_uint64(seed), _p(ncc), _p(tr)))
        return coord, normal, ncc, tr

    def post_process(self, coord, normal, ncc, views, nviews, maxv: Optional[int] = None):
        """Optim::postProcess (optim.cpp:260-290), store-independent part."""
        coord, normal, views, nviews = self._cv(coord, normal, views, nviews)
        ncc = np.ascontiguousarray(ncc, np.float32)
        n, stride = views.shape
        maxv = maxv or self.nviews
        ret, images, nimg = np.zeros(n, np.int32), np.zeros((n, maxv), np.int32), np.zeros(n, np.int32)
        grids, tmp = np.zeros((n, maxv, 2), np.int32), np.zeros(n, np.float32)
        _chk(lib().pmk_post_process(self.h, n, _p(coord), _p(normal), _p(ncc), _p(views), _p(nviews), stride, maxv, _p(ret), _p(images), _p(nimg), _p(grids), _p(tmp)))
        return ret, images, nimg, grids, tmp

    def probe(self, view, coord, normal=None):
        view = np.ascontiguousarray(view, np.int32)
        coord = np.ascontiguousarray(coord, np.float32)
        n = len(view)
        normal = np.ascontiguousarray(normal, np.float32) if normal is not None else None
        proj, unit = np.empty((n, 3), np.float32), np.empty(n, np.float32)
        px = np.empty((n, 4), np.float32) if normal is not None else None
        py = np.empty((n, 4), np.float32) if normal is not None else None
        ixy, ok = np.empty((n, 2), np.int32), np.empty(n, np.int32)
        _chk(lib().pmk_probe(self.h, n, _p(view), _p(coord), _p(normal), _p(proj), _p(unit), _p(px), _p(py), _p(ixy), _p(ok)))
        return dict(project=proj, unit=unit, px=px, py=py, cell=ixy, cell_ok=ok)

    def probe_unproject(self, view, icoord3) -> np.ndarray:
        """Camera::unproject (camera.cpp:329-337) at the working level."""
        view, icoord3 = np.ascontiguousarray(view, np.int32), np.ascontiguousarray(icoord3, np.float32)
        out = np.empty((len(view), 4), np.float32)
        _chk(lib().pmk_probe_unproject(self.h, len(view), _p(view), _p(icoord3), _p(out)))
        return out

    # -- device patch store, sweep and filters ----------------------------------------------------------------
    def store_clear(self):
        _chk(lib().pmk_store_clear(self.h))

    def store_add(self, coord, normal, scal, images, nimages):
        """PatchManager::readPatches body: setGrids + addPatch for each record."""
        coord, normal, images, nimages = self._cv(coord, normal, images, nimages)
        scal = np.ascontiguousarray(scal, np.float32)
        _chk(lib().pmk_store_add(self.h, len(coord), _p(coord), _p(normal), _p(scal), _p(images), _p(nimages), images.shape[1]))

    def store_count(self) -> int:
        n = C.c_int()
        _chk(lib().pmk_store_count(self.h, C.byref(n)))
        return n.value

    def store_get(self, maxv: Optional[int] = None) -> "Patches":
        n = self.store_count()
        out = Patches(n, maxv or self.nviews)
        got = C.c_int()
        _chk(lib().pmk_store_get(self.h, n, out.maxv, _p(out.coord), _p(out.normal), _p(out.scal), _p(out.images), _p(out.nimages), _p(out.grids),
                                 _p(out.vimages), _p(out.nvimages), _p(out.vgrids), C.byref(got)))
        assert got.value == n
        return out

    def store_depth_map(self, view: int) -> np.ndarray:
        gw, gh = self.grid_dims(view)
        ids = np.zeros((gh, gw), np.int32)
        _chk(lib().pmk_store_depth_map(self.h, view, _p(ids)))
        return ids

    def store_cell_counts(self, view: int, which: int = 0) -> np.ndarray:
        gw, gh = self.grid_dims(view)
        out = np.zeros((gh, gw), np.int32)
        _chk(lib().pmk_store_cell_counts(self.h, view, which, _p(out)))
        return out

    def store_colors(self, n: int) -> np.ndarray:
        out = np.zeros((n, 3), np.uint8)
        _chk(lib().pmk_store_colors(self.h, n, _p(out)))
        return out

    SWEEP_STATS = ("calls", "tries", "gen_null", "ncc_lose", "fail0", "fail1", "added", "replaced", "trimmed", "evals", "cell_ns", "step_max_ns", "steps", "nccl_ns", "msg_bytes")

    def propagate(self, it: int, seed: int) -> dict:
        """Propagate::run(iter)."""
        st = np.zeros(16, np.uint64)
        _chk(lib().pmk_propagate(self.h, it, C.c_uint64(seed), _p(st)))
        return dict(zip(self.SWEEP_STATS, (int(v) for v in st)))

    def propagate_diagonals(self, it: int, image: int, first: int, count: int, seed: int) -> dict:
        st = np.zeros(16, np.uint64)
        _chk(lib().pmk_propagate_diagonals(self.h, it, image, first, count, C.c_uint64(seed), _p(st)))
        return dict(zip(self.SWEEP_STATS, (int(v) for v in st)))

    def propagate_forced(self, it: int, image: int, x: int, y: int, code, ncc0, coord, normal, scal, images, nimages) -> dict:
        """pmk_propagate_forced: replay dest cell (x, y) with every try starting from a recorded post-refinePatch hypothesis."""
        code, nimages = np.ascontiguousarray(code, np.int32), np.ascontiguousarray(nimages, np.int32)
        ncc0 = np.ascontiguousarray(ncc0, np.float32)
        coord, normal, scal = (np.ascontiguousarray(a, np.float32) for a in (coord, normal, scal))
        images = np.ascontiguousarray(images, np.int32)
        T, S = len(code), images.shape[1]
        o = dict(ntries=np.zeros(1, np.int32), outcome=np.zeros(T, np.int32), branch_full=np.zeros(T, np.int32), post_ret=np.zeros(T, np.int32),
                 nimages=np.zeros(T, np.int32), images=np.zeros((T, S), np.int32), grids=np.zeros((T, S, 2), np.int32),
                 nvimages=np.zeros(T, np.int32), vimages=np.zeros((T, S), np.int32), vgrids=np.zeros((T, S, 2), np.int32), tmp=np.zeros(T, np.float32))
        io = ForcedIO(T, S, _p(code), _p(ncc0), _p(coord), _p(normal), _p(scal), _p(nimages), _p(images), _p(o["ntries"]), _p(o["outcome"]),
                      _p(o["branch_full"]), _p(o["post_ret"]), _p(o["nimages"]), _p(o["images"]), _p(o["grids"]), _p(o["nvimages"]), _p(o["vimages"]),
                      _p(o["vgrids"]), _p(o["tmp"]))
        st = np.zeros(16, np.uint64)
        _chk(lib().pmk_propagate_forced(self.h, it, image, x, y, C.byref(io), _p(st)))
        o["ntries"] = int(o["ntries"][0])
        o["stats"] = dict(zip(self.SWEEP_STATS, (int(v) for v in st)))
        return o

    def filter_rebuild(self, additive: int) -> int:
        n = C.c_int()
        _chk(lib().pmk_filter_rebuild(self.h, additive, C.byref(n)))
        return n.value

    def filter_stage(self, stage: int, n: int):
        """-> (f_out, i_out, i_out2, killed); see include/pmk.h for what each stage reports."""
        f, i1, i2 = np.zeros(n, np.float32), np.zeros(n, np.int32), np.zeros(n, np.int32)
        k = C.c_int()
        _chk(lib().pmk_filter_stage(self.h, stage, n, _p(f), _p(i1), _p(i2), C.byref(k)))
        return f, i1, i2, k.value

    def filter(self):
        """Filter::run -> [before, removed by outside / exact / neighbor / small groups, after]."""
        c = np.zeros(6, np.int32)
        _chk(lib().pmk_filter(self.h, _p(c)))
        return [int(v) for v in c]

    def probe_neighbor(self, lhs10, rhs10, thr: float, hunit=None, radius=None):
        """PmMvps::isNeighbor / isNeighborRadius on pairs; a patch = coord4, normal4, dscale, reference view."""
        lhs10, rhs10 = np.ascontiguousarray(lhs10, np.float32), np.ascontiguousarray(rhs10, np.float32)
        hunit = np.ascontiguousarray(hunit, np.float32) if hunit is not None else None
        radius = np.ascontiguousarray(radius, np.float32) if radius is not None else None
        out = np.zeros(len(lhs10), np.int32)
        _chk(lib().pmk_probe_neighbor(self.h, len(lhs10), _p(lhs10), _p(rhs10), _p(hunit), _p(radius), C.c_float(thr), _p(out)))
        return out

    def probe_check(self, coord, normal, scal, images, nimages):
        """setVImagesVGrids + Optim::check on free-standing candidates -> ret, gain, neighbour count, vimages, nvimages."""
        coord, normal, images, nimages = self._cv(coord, normal, images, nimages)
        scal = np.ascontiguousarray(scal, np.float32)
        n, stride = images.shape
        ret, gain, nn = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros(n, np.int32)
        vimg, nvimg = np.zeros((n, stride), np.int32), np.zeros(n, np.int32)
        _chk(lib().pmk_probe_check(self.h, n, _p(coord), _p(normal), _p(scal), _p(images), _p(nimages), stride, _p(ret), _p(gain), _p(nn), _p(vimg), _p(nvimg)))
        return ret, gain, nn, vimg, nvimg

    # -- PatchManager pass-throughs ----------------------------------------------------------------------------
    def probe_visible(self, coord, normal, image, cells=None, strict: float = 0.5):
        """PatchManager::isVisible0 (cells None) / isVisible (patch_manager.cpp:327-376) -> (flags, cells)."""
        coord, normal = np.ascontiguousarray(coord, np.float32), np.ascontiguousarray(normal, np.float32)
        image = np.ascontiguousarray(image, np.int32)
        cells = np.ascontiguousarray(cells, np.int32) if cells is not None else None
        out, cout = np.zeros(len(image), np.int32), np.zeros((len(image), 2), np.int32)
        _chk(lib().pmk_probe_visible(self.h, len(image), _p(coord), _p(normal), _p(image), _p(cells), C.c_float(strict), _p(out), _p(cout)))
        return out, cout

    def probe_scales(self, coord, images, nimages):
        """PatchManager::setScales (patch_manager.cpp:378-399) -> m_dscale, m_ascale."""
        coord, images, nimages = np.ascontiguousarray(coord, np.float32), np.ascontiguousarray(images, np.int32), np.ascontiguousarray(nimages, np.int32)
        ds, asc = np.zeros(len(coord), np.float32), np.zeros(len(coord), np.float32)
        _chk(lib().pmk_probe_scales(self.h, len(coord), _p(coord), _p(images), _p(nimages), images.shape[1], _p(ds), _p(asc)))
        return ds, asc

    def probe_neighbors(self, coord, normal, scal, images, nimages, scale: float = 4.0, margin: int = 2, cap: int = 512):
        """PatchManager::findNeighbors (patch_manager.cpp:671-728) -> list of ascending id arrays."""
        coord, normal, scal = (np.ascontiguousarray(a, np.float32) for a in (coord, normal, scal))
        images, nimages = np.ascontiguousarray(images, np.int32), np.ascontiguousarray(nimages, np.int32)
        n = len(coord)
        ids, cnt = np.zeros((n, cap), np.int32), np.zeros(n, np.int32)
        _chk(lib().pmk_probe_neighbors(self.h, n, _p(coord), _p(normal), _p(scal), _p(images), _p(nimages), images.shape[1], C.c_float(scale), margin, cap, _p(ids), _p(cnt)))
        return [ids[i, :min(cnt[i], cap)].copy() for i in range(n)], cnt

    def store_ids(self):
        """Store ids of the live patches in collect order (row i of store_get() is store id ids[i]); arange(n) right after a rebuild."""
        n = self.store_count()
        ids = np.zeros(max(n, 1), np.int32)
        got = C.c_int()
        _chk(lib().pmk_store_ids(self.h, n, _p(ids), C.byref(got)))
        assert got.value == n
        return ids[:n]

    def store_remove(self, ids):
        ids = np.ascontiguousarray(ids, np.int32)
        _chk(lib().pmk_store_remove(self.h, len(ids), _p(ids)))

    def store_update_depth_maps(self, ids):
        ids = np.ascontiguousarray(ids, np.int32)
        _chk(lib().pmk_store_update_depth_maps(self.h, len(ids), _p(ids)))

    def store_cell_ids(self, view: int, which: int = 0):
        """m_pgrids (which 0) / m_vpgrids (1) of one view: (offsets[cells + 1], ids)."""
        gw, gh = self.grid_dims(view)
        offs, tot = np.zeros(gw * gh + 1, np.int32), C.c_int()
        _chk(lib().pmk_store_cell_ids(self.h, view, which, _p(offs), None, 0, C.byref(tot)))
        ids = np.zeros(max(tot.value, 1), np.int32)
        _chk(lib().pmk_store_cell_ids(self.h, view, which, _p(offs), _p(ids), len(ids), C.byref(tot)))
        return offs, ids[:tot.value]

    # -- multi-GPU --------------------------------------------------------------------------------------------
    def comm_init(self, rank: int, nranks: int, unique_id: Optional[bytes]):
        _chk(lib().pmk_comm_init(self.h, rank, nranks, unique_id))

    def store_checksum(self):
        out = np.zeros(2, np.uint64)
        _chk(lib().pmk_store_checksum(self.h, _p(out)))
        return int(out[0]), int(out[1])

    # -- plumbing ------------------------------------------------------------------------------------
    def alloc(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    def sync(self):
        _chk(lib().pmk_sync(self.h))

    def timer_begin(self):
        _chk(lib().pmk_timer_begin(self.h))

    def timer_end(self) -> float:
        ms = C.c_float()
        _chk(lib().pmk_timer_end(self.h, C.byref(ms)))
        return float(ms.value)

    def launch_count(self) -> int:
        v = C.c_uint64()
        _chk(lib().pmk_launch_count(self.h, C.byref(v)))
        return int(v.value)

    def flush_l2(self):
        _chk(lib().pmk_flush_l2(self.h))


def contour2_to_projection(intrinsics6, extrinsics6) -> np.ndarray:
    """Camera::setProjection for CONTOUR2 camera files (camera.cpp:116-131, 241-261) -> level-0 P (3, 4)."""
    a, b = np.ascontiguousarray(intrinsics6, np.float32), np.ascontiguousarray(extrinsics6, np.float32)
    P = np.zeros((3, 4), np.float32)
    _chk(lib().pmk_contour2_to_projection(_p(a), _p(b), _p(P)))
    return P


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array over cudaMallocHost memory (kept alive for the life of the process)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = C.c_void_p()
    _chk(lib().pmk_host_alloc_pinned(C.c_uint64(max(n, 1)), C.byref(ptr)))
    buf = (C.c_char * max(n, 1)).from_address(ptr.value)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def exported_symbols():
    """Names declared in include/pmk.h, for the CPU-side symbol test."""
    import re
    hdr = os.path.join(HERE, "..", "include", "pmk.h")
    txt = open(hdr).read()
    return sorted(set(re.findall(r"\b(pmk_[a-z0-9_]+)\s*\(", txt)))


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _chk(lib().pmk_comm_unique_id(buf))
    return buf.raw


def step_share(step_tasks: int, rank: int, nranks: int) -> int:
    """How many of a wavefront step's dest cells rank `rank` of `nranks` sweeps (pmk_step_share)."""
    n = C.c_int()
    _chk(lib().pmk_step_share(step_tasks, rank, nranks, C.byref(n)))
    return n.value
