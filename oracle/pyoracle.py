"""oracle/pyoracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes bindings for the two CPU checkers:

  * ``RefLib``  -> oracle/_ref/libpmref.so   : the reference's own sources (built by oracle/Makefile
                                               from /root/reference; prebuilt file travels to the GPU box)
  * ``COracle`` -> oracle/_ref/libpmoracle.so: the plain-C restatement (oracle/pm_oracle.c)

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg / --impl reference) import this.
The product package ``mvskit_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libpmref.so")
ORACLE_SO = os.path.join(HERE, "_ref", "libpmoracle.so")


def build(ref: bool = True) -> None:
    """Compile the C restatement, and the reference itself when /root/reference is present."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir("/root/reference/pmmvps"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def _i32(a):
    return np.ascontiguousarray(a, np.int32)


class PatchIO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("coord4", "normal4", "scal4", "images", "nimages", "grids", "vimages", "nvimages", "vgrids")] + [("maxv", C.c_int)]


class PatchBatch:
    """numpy-side mirror of the harness' PatchIO record arrays."""

    def __init__(self, n: int, maxv: int):
        self.n, self.maxv = n, maxv
        self.coord = np.zeros((n, 4), np.float32)
        self.normal = np.zeros((n, 4), np.float32)
        self.scal = np.zeros((n, 4), np.float32)      # ncc, dscale, ascale, tmp
        self.images = np.full((n, maxv), -1, np.int32)
        self.nimages = np.zeros(n, np.int32)
        self.grids = np.zeros((n, maxv, 2), np.int32)
        self.vimages = np.full((n, maxv), -1, np.int32)
        self.nvimages = np.zeros(n, np.int32)
        self.vgrids = np.zeros((n, maxv, 2), np.int32)

    def io(self) -> PatchIO:
        return PatchIO(_p(self.coord), _p(self.normal), _p(self.scal), _p(self.images), _p(self.nimages), _p(self.grids),
                       _p(self.vimages), _p(self.nvimages), _p(self.vgrids), self.maxv)


class TraceIO(C.Structure):
    _fields_ = [("cap", C.c_int), ("code", C.c_void_p), ("ncc0", C.c_void_p), ("post_ret", C.c_void_p), ("decision", C.c_void_p),
                ("branch_full", C.c_void_p), ("mid", PatchIO), ("fin", PatchIO)]


class Trace:
    """Per-try record of pmref_trace_dest (oracle/ref_harness.cpp): code 0 generatePatch NULL / 1 lost to the worst patch /
    2 preProcess == -1 / 3 refined; mid = the patch as refinePatch left it, fin = as postProcess left it; decision 0 / 1 added /
    2 replaced the worst."""

    def __init__(self, cap: int, maxv: int):
        self.cap = cap
        self.code = np.zeros(cap, np.int32)
        self.ncc0 = np.zeros(cap, np.float32)
        self.post_ret = np.zeros(cap, np.int32)
        self.decision = np.zeros(cap, np.int32)
        self.branch_full = np.zeros(cap, np.int32)
        self.mid = PatchBatch(cap, maxv)
        self.fin = PatchBatch(cap, maxv)
        self.n = 0

    def io(self) -> TraceIO:
        return TraceIO(self.cap, _p(self.code), _p(self.ncc0), _p(self.post_ret), _p(self.decision), _p(self.branch_full), self.mid.io(), self.fin.io())


class RefLib:
    """The compiled reference.  One scene per process (the reference keeps a static Optim::m_inst)."""

    def __init__(self, prefix: str, option: str = "option"):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO + " (run `make -C oracle ref` where /root/reference exists)")
        self.L = L = C.CDLL(REF_SO)
        L.pmref_time_compute_ncc.restype = C.c_double
        L.pmref_run.restype = C.c_double
        L.pmref_threshold.restype = C.c_float
        if not prefix.endswith("/"):
            prefix += "/"
        L.pmref_init(prefix.encode(), option.encode())
        self.nviews, self.level, self.csize, self.wsize, self.tau, self.min_image_num = (L.pmref_info(i) for i in range(6))
        self.nlevels = L.pmref_info(7)
        self.init_thresholds = [float(L.pmref_threshold(i)) for i in range(9)]      # as PmMvps::init left them (pmmvps.cpp:52-67)

    # -- scalars ---------------------------------------------------------------------------------
    def threshold(self, what: int) -> float:
        return float(self.L.pmref_threshold(what))

    def set_depth(self, d: int):
        self.L.pmref_set_depth(d)

    def update_threshold(self):
        self.L.pmref_update_threshold()

    def set_ncc_thresholds(self, ncc: float, before: float):
        self.L.pmref_set_ncc_thresholds(C.c_float(ncc), C.c_float(before))

    def log_is_double(self) -> bool:
        return bool(self.L.pmref_log_is_double())

    def level_diff(self, ratio: float) -> int:
        return int(self.L.pmref_level_diff(C.c_float(ratio)))

    # -- scene -----------------------------------------------------------------------------------
    def image_dims(self, view: int, level: int):
        w, h = C.c_int(), C.c_int()
        self.L.pmref_image_dims(view, level, C.byref(w), C.byref(h))
        return w.value, h.value

    def grid_dims(self, view: int):
        w, h = C.c_int(), C.c_int()
        self.L.pmref_grid_dims(view, C.byref(w), C.byref(h))
        return w.value, h.value

    def image(self, view: int, level: int) -> np.ndarray:
        w, h = self.image_dims(view, level)
        out = np.zeros((h, w, 3), np.uint8)
        self.L.pmref_get_image(view, level, _p(out))
        return out

    def camera(self, view: int, level: int = 0):
        P, c, o = np.zeros(12, np.float32), np.zeros(4, np.float32), np.zeros(4, np.float32)
        x, y, z, ip = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32), C.c_float()
        self.L.pmref_get_camera(view, level, _p(P), _p(c), _p(o), _p(x), _p(y), _p(z), C.byref(ip))
        return dict(P=P.reshape(3, 4), center=c, oaxis=o, xaxis=x, yaxis=y, zaxis=z, ipscale=np.float32(ip.value))

    # -- per-function probes -----------------------------------------------------------------------
    def mask_level(self, view: int, level: int):
        """Image::m_masks[level] of one view (None when the view has no mask)."""
        w, h = self.image_dims(view, level)
        out = np.zeros((h, w), np.uint8)
        return out if self.L.pmref_get_mask_level(view, level, _p(out)) else None

    def get_mask(self, coord, view: int = -1, level=None) -> np.ndarray:
        """PhotoSet::getMask(coord, level) (view < 0) or PhotoSet::getMask(view, coord, level)."""
        coord = _f32(coord)
        out = np.zeros(len(coord), np.int32)
        self.L.pmref_get_mask(len(coord), view, _p(coord), self.level if level is None else level, _p(out))
        return out

    def project(self, views, coord, level=None):
        views, coord = _i32(views), _f32(coord)
        out = np.zeros((len(views), 3), np.float32)
        self.L.pmref_project(len(views), _p(views), _p(coord), self.level if level is None else level, _p(out))
        return out

    def unproject(self, views, icoord, level=None):
        views, icoord = _i32(views), _f32(icoord)
        out = np.zeros((len(views), 4), np.float32)
        self.L.pmref_unproject(len(views), _p(views), _p(icoord), self.level if level is None else level, _p(out))
        return out

    def get_unit(self, views, coord):
        views, coord = _i32(views), _f32(coord)
        out = np.zeros(len(views), np.float32)
        self.L.pmref_get_unit(len(views), _p(views), _p(coord), _p(out))
        return out

    def get_paxes(self, views, coord, normal):
        views, coord, normal = _i32(views), _f32(coord), _f32(normal)
        px, py = np.zeros((len(views), 4), np.float32), np.zeros((len(views), 4), np.float32)
        self.L.pmref_get_paxes(len(views), _p(views), _p(coord), _p(normal), _p(px), _p(py))
        return px, py

    def get_color(self, views, xy, level):
        views, xy = _i32(views), _f32(xy)
        out = np.zeros((len(views), 3), np.float32)
        self.L.pmref_get_color(len(views), _p(views), _p(xy), level, _p(out))
        return out

    def cells(self, views, coord):
        views, coord = _i32(views), _f32(coord)
        ixy, ok = np.zeros((len(views), 2), np.int32), np.zeros(len(views), np.int32)
        self.L.pmref_cells(len(views), _p(views), _p(coord), _p(ixy), _p(ok))
        return ixy, ok

    def get_tex(self, coord, normal, refview, view):
        tex = np.zeros((self.wsize * self.wsize, 3), np.float32)
        flag, level = C.c_int(), C.c_int()
        self.L.pmref_get_tex(_p(_f32(coord)), _p(_f32(normal)), refview, view, _p(tex), C.byref(flag), C.byref(level))
        return tex, flag.value, (level.value if flag.value == 0 else -1)

    def weights(self, coord, normal, views):
        views = _i32(views)
        w = np.zeros(len(views), np.float32)
        self.L.pmref_weights(_p(_f32(coord)), _p(_f32(normal)), _p(views), len(views), _p(w))
        return w

    def compute_ncc(self, coord, normal, views, nviews):
        coord, normal, views, nviews = _f32(coord), _f32(normal), _i32(views), _i32(nviews)
        n = len(coord)
        incc, ncc = np.zeros(n, np.float32), np.zeros(n, np.float32)
        self.L.pmref_compute_ncc(n, _p(coord), _p(normal), _p(views), _p(nviews), views.shape[1], _p(incc), _p(ncc))
        return incc, ncc

    def time_compute_ncc(self, coord, normal, views, nviews, repeats: int = 1) -> float:
        coord, normal, views, nviews = _f32(coord), _f32(normal), _i32(views), _i32(nviews)
        chk = C.c_double()
        return float(self.L.pmref_time_compute_ncc(len(coord), _p(coord), _p(normal), _p(views), _p(nviews), views.shape[1], repeats, C.byref(chk)))

    def set_inccs(self, coord, normal, views, robust: int):
        views = _i32(views)
        out = np.zeros(len(views), np.float32)
        self.L.pmref_set_inccs(_p(_f32(coord)), _p(_f32(normal)), _p(views), len(views), robust, _p(out))
        return out

    def set_inccs_pair(self, coord, normal, views, robust: int):
        views = _i32(views)
        out = np.zeros((len(views), len(views)), np.float32)
        self.L.pmref_set_inccs_pair(_p(_f32(coord)), _p(_f32(normal)), _p(views), len(views), robust, _p(out))
        return out

    # -- multi-step ------------------------------------------------------------------------------------
    def pre_process(self, coord, normal, views, nviews, maxv=None):
        coord, normal, views, nviews = _f32(coord), _f32(normal), _i32(views), _i32(nviews)
        n = len(coord)
        out = PatchBatch(n, maxv or self.nviews)
        ret = np.zeros(n, np.int32)
        io = out.io()
        self.L.pmref_pre_process(n, _p(coord), _p(normal), _p(views), _p(nviews), views.shape[1], _p(ret), C.byref(io))
        return ret, out

    def post_process(self, coord, normal, scal, views, nviews, maxv=None):
        coord, normal, scal, views, nviews = _f32(coord), _f32(normal), _f32(scal), _i32(views), _i32(nviews)
        n = len(coord)
        out = PatchBatch(n, maxv or self.nviews)
        ret = np.zeros(n, np.int32)
        io = out.io()
        self.L.pmref_post_process(n, _p(coord), _p(normal), _p(scal), _p(views), _p(nviews), views.shape[1], _p(ret), C.byref(io))
        return ret, out

    def refine(self, coord, normal, dscale, views, nviews, streams, seed: int, trace: bool = False):
        coord, normal, dscale = _f32(coord).copy(), _f32(normal).copy(), _f32(dscale)
        views, nviews = _i32(views), _i32(nviews)
        streams = np.ascontiguousarray(streams, np.uint64)
        n = len(coord)
        ncc = np.zeros(n, np.float32)
        tr = np.zeros((n, 97, 4), np.float64) if trace else None
        self.L.pmref_refine_seed(C.c_ulonglong(seed))
        self.L.pmref_refine(n, _p(coord), _p(normal), _p(dscale), _p(views), _p(nviews), views.shape[1], _p(streams), _p(ncc), _p(tr))
        return coord, normal, ncc, tr

    def cost_func(self, coord, normal, dscale, views, x):
        views, x = _i32(views), np.ascontiguousarray(x, np.float64)
        out = np.zeros(len(x), np.float64)
        self.L.pmref_cost_func(_p(_f32(coord)), _p(_f32(normal)), C.c_float(dscale), _p(views), len(views), len(x), _p(x), _p(out))
        return out

    def encode(self, coord, normal, dscale, refview):
        x = np.zeros(3, np.float64)
        self.L.pmref_encode(_p(_f32(coord)), _p(_f32(normal)), C.c_float(dscale), refview, _p(x))
        return x

    # -- patch store ---------------------------------------------------------------------------------
    def clear_patches(self):
        self.L.pmref_clear_patches()

    def add_patches(self, coord, normal, scal, views, nviews):
        coord, normal, scal, views, nviews = _f32(coord), _f32(normal), _f32(scal), _i32(views), _i32(nviews)
        self.L.pmref_add_patches(len(coord), _p(coord), _p(normal), _p(scal), _p(views), _p(nviews), views.shape[1])

    def create_patches(self):
        self.L.pmref_create_patches()

    def collect(self, target: int = 0) -> int:
        return int(self.L.pmref_collect(target))

    def get_patches(self, maxv=None) -> PatchBatch:
        n = self.collect(0)
        out = PatchBatch(n, maxv or self.nviews)
        io = out.io()
        self.L.pmref_get_patches(C.byref(io))
        return out

    def depth_map(self, view: int) -> np.ndarray:
        gw, gh = self.grid_dims(view)
        ids = np.zeros((gh, gw), np.int32)
        self.L.pmref_get_depth_map(view, _p(ids))
        return ids

    def cell_counts(self, view: int, which: int = 0) -> np.ndarray:
        gw, gh = self.grid_dims(view)
        out = np.zeros((gh, gw), np.int32)
        self.L.pmref_get_cell_counts(view, which, _p(out))
        return out

    def propagate_run(self, it: int):
        self.L.pmref_propagate_run(it)

    def filter_run(self):
        self.L.pmref_filter_run()

    def gains(self) -> np.ndarray:
        n = self.collect(1)
        out = np.zeros(n, np.float32)
        self.L.pmref_gains(_p(out))
        return out

    def is_neighbor(self, a, b, thr: float):
        a, b = _i32(a), _i32(b)
        out = np.zeros(len(a), np.int32)
        self.L.pmref_is_neighbor(len(a), _p(a), _p(b), C.c_float(thr), _p(out))
        return out

    # -- schedule PMS1 through the reference's propagatePatch; Filter::run stage by stage -----------------
    def propagate_dest(self, image: int, x: int, y: int, inc: int, it: int) -> int:
        return int(self.L.pmref_propagate_dest(image, x, y, inc, it))

    def trace_dest(self, image: int, x: int, y: int, inc: int, it: int, cap: int = 64) -> Trace:
        """propagate_dest with every try recorded (teacher forcing); the store ends up exactly as after propagate_dest."""
        tr = Trace(cap, self.nviews)
        io = tr.io()
        tr.n = int(self.L.pmref_trace_dest(image, x, y, inc, it, C.byref(io)))
        assert tr.n <= cap, (tr.n, cap)
        return tr

    def propagate_diag(self, image: int, diag: int, inc: int, it: int) -> int:
        return int(self.L.pmref_propagate_diag(image, diag, inc, it))

    def refine_seed(self, seed: int):
        self.L.pmref_refine_seed(C.c_ulonglong(seed))

    def cell_ids(self, view: int, index: int, which: int = 0, cap: int = 256) -> np.ndarray:
        ids = np.zeros(cap, np.int32)
        n = int(self.L.pmref_cell_ids(view, index, which, _p(ids), cap))
        return ids[:min(n, cap)]

    def filter_rebuild(self, additive: int):
        self.L.pmref_filter_rebuild(additive)

    def stage_begin(self) -> int:
        self._nstage = int(self.L.pmref_stage_begin())
        return self._nstage

    def filter_stage(self, stage: int):
        self.L.pmref_filter_stage(stage)

    def stage_alive(self) -> np.ndarray:
        out = np.zeros(self._nstage, np.int32)
        self.L.pmref_stage_alive(_p(out))
        return out

    def stage_patches(self, maxv=None) -> PatchBatch:
        out = PatchBatch(self._nstage, maxv or self.nviews)
        io = out.io()
        self.L.pmref_stage_patches(C.byref(io))
        return out

    def stage_gains(self) -> np.ndarray:
        out = np.zeros(self._nstage, np.float32)
        self.L.pmref_stage_gains(_p(out))
        return out

    def stage_rejects(self) -> np.ndarray:
        out = np.zeros(self._nstage, np.int32)
        self.L.pmref_stage_rejects(_p(out))
        return out

    def stage_neighbors(self):
        cnt, quad = np.zeros(self._nstage, np.int32), np.zeros(self._nstage, np.int32)
        self.L.pmref_stage_neighbors(_p(cnt), _p(quad))
        return cnt, quad

    def check(self, coord, normal, scal, views):
        coord, normal, scal, views = _f32(coord), _f32(normal), _f32(scal), _i32(views)
        out = PatchBatch(1, self.nviews)
        io = out.io()
        gain = C.c_float()
        r = int(self.L.pmref_check(_p(coord), _p(normal), _p(scal), _p(views), len(views), C.byref(gain), C.byref(io)))
        return r, float(gain.value), out

    def write_ply(self, path: str) -> int:
        """PatchManager::writePly of the current store."""
        return int(self.L.pmref_write_ply(path.encode()))

    # -- PatchManager's public surface --------------------------------------------------------------
    def is_visible(self, coord, normal, image, cells=None, strict: float = 0.5):
        coord, normal, image = _f32(coord), _f32(normal), _i32(image)
        cells = _i32(cells) if cells is not None else None
        out, cout = np.zeros(len(image), np.int32), np.zeros((len(image), 2), np.int32)
        self.L.pmref_is_visible(len(image), _p(coord), _p(normal), _p(image), _p(cells) if cells is not None else None, C.c_float(strict), _p(out), _p(cout))
        return out, cout

    def set_scales(self, coord, views, nviews):
        coord, views, nviews = _f32(coord), _i32(views), _i32(nviews)
        ds, asc = np.zeros(len(coord), np.float32), np.zeros(len(coord), np.float32)
        self.L.pmref_set_scales(len(coord), _p(coord), _p(views), _p(nviews), views.shape[1], _p(ds), _p(asc))
        return ds, asc

    def find_neighbors(self, coord, normal, scal, views, scale: float = 4.0, margin: int = 2, cap: int = 4096):
        views = _i32(views)
        ids = np.zeros(cap, np.int32)
        n = int(self.L.pmref_find_neighbors(_p(_f32(coord)), _p(_f32(normal)), _p(_f32(scal)), _p(views), len(views), C.c_float(scale), margin, _p(ids), cap))
        return ids[:min(n, cap)].copy(), n

    def remove_patches(self, ids):
        ids = _i32(ids)
        self.L.pmref_remove_patches(len(ids), _p(ids))

    def update_depth_maps(self, ids):
        ids = _i32(ids)
        self.L.pmref_update_depth_maps(len(ids), _p(ids))

    def run(self):
        alive = C.c_int()
        secs = float(self.L.pmref_run(C.byref(alive)))
        return secs, alive.value


class COracle:
    """The plain-C restatement, fed with arrays (no files)."""

    def __init__(self, P: np.ndarray, images, level=1, csize=2, wsize=7, min_image_num=3, ncc_threshold=0.7, masks=None):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        self.L = L = C.CDLL(ORACLE_SO)
        L.pmo_scene_create.restype = C.c_void_p
        L.pmo_get_unit.restype = C.c_float
        L.pmo_compute_incc.restype = C.c_float
        n = len(images)
        self.nviews, self.level, self.csize, self.wsize = n, level, csize, wsize
        self.tau = min(2 * min_image_num, n)
        self.nlevels = level + 3
        self.s = C.c_void_p(L.pmo_scene_create(n, level, csize, wsize, min_image_num, C.c_float(ncc_threshold)))
        P = _f32(P).reshape(n, 12)
        for v in range(n):
            L.pmo_set_camera(self.s, v, _p(P[v]))
            im = np.ascontiguousarray(images[v], np.uint8)
            L.pmo_set_image(self.s, v, _p(im), im.shape[1], im.shape[0])
            if masks and masks[v] is not None:
                m = np.ascontiguousarray(masks[v], np.uint8)
                L.pmo_set_mask(self.s, v, _p(m), m.shape[1], m.shape[0])

    def mask_level(self, view: int, level: int):
        S = self._scene()
        w, h = self.image_dims(view, level)
        ptrs = np.ctypeslib.as_array(C.cast(S.mask, C.POINTER(C.c_uint64)), (self.nviews * self.nlevels,))
        p = int(ptrs[view * self.nlevels + level])
        return self._arr(p, (h, w), np.uint8) if p else None

    def get_mask(self, coord, view: int = -1, level=None) -> np.ndarray:
        coord = _f32(coord)
        lv = self.level if level is None else level
        out = np.zeros(len(coord), np.int32)
        for i in range(len(coord)):
            out[i] = self.L.pmo_get_mask(self.s, _p(coord[i]), lv) if view < 0 else self.L.pmo_get_mask_view(self.s, view, _p(coord[i]), lv)
        return out

    def __del__(self):
        try:
            self.L.pmo_scene_destroy(self.s)
        except Exception:
            pass

    def _scene(self):
        class S(C.Structure):
            _fields_ = [(k, C.c_int) for k in ("nviews", "level", "nlevels", "csize", "wsize", "tau", "min_image_num", "depth")] + \
                       [(k, C.c_float) for k in ("ncc_threshold", "ncc_threshold_before", "angle_threshold0", "angle_threshold1",
                                                 "max_angle_threshold", "quad_threshold", "neighbor_threshold", "neighbor_threshold1", "neighbor_threshold2")] + \
                       [(k, C.c_void_p) for k in ("P", "center", "oaxis", "xaxis", "yaxis", "zaxis", "ipscale", "Minv", "img", "w", "h", "gw", "gh", "mask")]
        return C.cast(self.s, C.POINTER(S)).contents

    def _arr(self, ptr, shape, dtype):
        n = int(np.prod(shape))
        ct = {np.float32: C.c_float, np.int32: C.c_int, np.uint8: C.c_ubyte}[dtype]
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), (n,)).reshape(shape).copy()

    def camera(self, view: int, level: int = 0):
        S = self._scene()
        n, nl = self.nviews, self.nlevels
        return dict(P=self._arr(S.P, (n, nl, 3, 4), np.float32)[view, level], center=self._arr(S.center, (n, 4), np.float32)[view],
                    oaxis=self._arr(S.oaxis, (n, 4), np.float32)[view], xaxis=self._arr(S.xaxis, (n, 3), np.float32)[view],
                    yaxis=self._arr(S.yaxis, (n, 3), np.float32)[view], zaxis=self._arr(S.zaxis, (n, 3), np.float32)[view],
                    ipscale=self._arr(S.ipscale, (n,), np.float32)[view])

    def image_dims(self, view: int, level: int):
        S = self._scene()
        n, nl = self.nviews, self.nlevels
        return int(self._arr(S.w, (n, nl), np.int32)[view, level]), int(self._arr(S.h, (n, nl), np.int32)[view, level])

    def image(self, view: int, level: int) -> np.ndarray:
        S = self._scene()
        w, h = self.image_dims(view, level)
        ptrs = np.ctypeslib.as_array(C.cast(S.img, C.POINTER(C.c_uint64)), (self.nviews * self.nlevels,))
        return self._arr(int(ptrs[view * self.nlevels + level]), (h, w, 3), np.uint8)

    def project(self, views, coord, level=None):
        views, coord = _i32(views), _f32(coord)
        out = np.zeros((len(views), 3), np.float32)
        for i in range(len(views)):
            self.L.pmo_project(self.s, int(views[i]), _p(coord[i]), self.level if level is None else level, _p(out[i]))
        return out

    def unproject(self, views, icoord, level=None):
        views, icoord = _i32(views), _f32(icoord)
        out = np.zeros((len(views), 4), np.float32)
        for i in range(len(views)):
            self.L.pmo_unproject(self.s, int(views[i]), _p(icoord[i]), self.level if level is None else level, _p(out[i]))
        return out

    def get_unit(self, views, coord):
        views, coord = _i32(views), _f32(coord)
        return np.array([self.L.pmo_get_unit(self.s, int(views[i]), _p(coord[i])) for i in range(len(views))], np.float32)

    def get_paxes(self, views, coord, normal):
        views, coord, normal = _i32(views), _f32(coord), _f32(normal)
        px, py = np.zeros((len(views), 4), np.float32), np.zeros((len(views), 4), np.float32)
        for i in range(len(views)):
            self.L.pmo_get_paxes(self.s, int(views[i]), _p(coord[i]), _p(normal[i]), _p(px[i]), _p(py[i]))
        return px, py

    def get_color(self, views, xy, level):
        views, xy = _i32(views), _f32(xy)
        out = np.zeros((len(views), 3), np.float32)
        for i in range(len(views)):
            self.L.pmo_get_color(self.s, int(views[i]), C.c_float(xy[i, 0]), C.c_float(xy[i, 1]), level, _p(out[i]))
        return out

    def cells(self, views, coord):
        views, coord = _i32(views), _f32(coord)
        ixy, ok = np.zeros((len(views), 2), np.int32), np.zeros(len(views), np.int32)
        for i in range(len(views)):
            ix, iy = C.c_int(), C.c_int()
            ok[i] = self.L.pmo_cell(self.s, int(views[i]), _p(coord[i]), C.byref(ix), C.byref(iy))
            ixy[i] = (ix.value, iy.value)
        return ixy, ok

    def level_diff(self, ratio: float) -> int:
        return int(self.L.pmo_level_diff(C.c_float(ratio)))

    def get_tex(self, coord, normal, refview, view):
        coord, normal = _f32(coord), _f32(normal)
        px, py = np.zeros(4, np.float32), np.zeros(4, np.float32)
        self.L.pmo_get_paxes(self.s, refview, _p(coord), _p(normal), _p(px), _p(py))
        tex = np.zeros((self.wsize * self.wsize, 3), np.float32)
        lvl = C.c_int()
        flag = self.L.pmo_get_tex(self.s, _p(coord), _p(px), _p(py), _p(normal), view, _p(tex), C.byref(lvl))
        if flag != 0:
            tex[:] = 0
        return tex, flag, lvl.value

    def weights(self, coord, normal, views):
        views = _i32(views)
        w = np.zeros(len(views), np.float32)
        self.L.pmo_compute_weights(self.s, _p(_f32(coord)), _p(_f32(normal)), _p(views), len(views), _p(w))
        return w

    def compute_ncc(self, coord, normal, views, nviews, want_levels: bool = False):
        coord, normal, views, nviews = _f32(coord), _f32(normal), _i32(views), _i32(nviews)
        n = len(coord)
        incc, ncc = np.zeros(n, np.float32), np.zeros(n, np.float32)
        levels = np.full((n, self.tau), -1, np.int32) if want_levels else None
        self.L.pmo_compute_ncc(self.s, n, _p(coord), _p(normal), _p(views), _p(nviews), views.shape[1], _p(incc), _p(ncc), _p(levels))
        return (incc, ncc, levels) if want_levels else (incc, ncc)

    def set_inccs(self, coord, normal, views, robust: int):
        views = _i32(views)
        out = np.zeros(len(views), np.float32)
        self.L.pmo_set_inccs(self.s, _p(_f32(coord)), _p(_f32(normal)), _p(views), len(views), robust, _p(out))
        return out

    def set_inccs_pair(self, coord, normal, views, robust: int):
        views = _i32(views)
        out = np.zeros((len(views), len(views)), np.float32)
        self.L.pmo_set_inccs_pair(self.s, _p(_f32(coord)), _p(_f32(normal)), _p(views), len(views), robust, _p(out))
        return out
