"""CPU-side checks of the drop-in boundary: the library loads and exports every symbol pmk.h declares."""
import ctypes
import os

import pytest

from conftest import HAVE_GPU


def test_library_exports_every_declared_symbol():
    from mvskit_b200 import build, pmk
    build.build()
    L = pmk.lib()
    names = pmk.exported_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.pmk_abi_version() == 3


def test_no_cpu_fallback():
    """Without a device the product must fail loudly, not compute on the CPU."""
    if HAVE_GPU:
        pytest.skip("a GPU is present")
    from mvskit_b200 import pmk
    with pytest.raises(pmk.PmkError, match="no CUDA device"):
        pmk.Context(nviews=5)


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "mvskit_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "pm_oracle" not in txt and "libpmref" not in txt.replace("oracle/_ref/libpmref.so, whose", ""), os.path.join(dirpath, f)


def test_bad_arguments():
    from mvskit_b200 import pmk
    L = pmk.lib()
    cfg = pmk.Config()
    L.pmk_default_config(ctypes.byref(cfg))
    assert (cfg.level, cfg.csize, cfg.wsize, cfg.min_image_num) == (1, 2, 7, 3)      # option.cpp:19-33
    h = ctypes.c_void_p()
    cfg.nviews = 0
    assert L.pmk_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    cfg.nviews, cfg.wsize = 5, 8
    assert L.pmk_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"wsize" in L.pmk_last_error()


def test_pack_hypotheses_layout():
    """The byte-lean records of pmk_ncc_eval_packed: xyz of coord / normal, byte view ids with 255 for "no such view", byte counts."""
    import numpy as np
    from mvskit_b200 import pmk
    coord = np.array([[1, 2, 3, 1], [4, 5, 6, 1]], np.float32)
    normal = np.array([[0, 0, 1, 0], [0, 1, 0, 0]], np.float32)
    views = np.array([[3, 0, 254, -1], [7, -1, -1, 300]], np.int32)
    nviews = np.array([3, 1], np.int32)
    c3, n3, v8, nv8 = pmk.Context.pack_hypotheses(coord, normal, views, nviews)
    assert c3.dtype == np.float32 and c3.shape == (2, 3) and c3.flags["C_CONTIGUOUS"] and np.array_equal(c3, coord[:, :3])
    assert np.array_equal(n3, normal[:, :3])
    assert v8.dtype == np.uint8 and v8.tolist() == [[3, 0, 254, 255], [7, 255, 255, 255]]
    assert nv8.dtype == np.uint8 and nv8.tolist() == [3, 1]
    assert c3.nbytes + n3.nbytes + v8.nbytes + nv8.nbytes == 2 * (12 + 12 + 4 + 1)
