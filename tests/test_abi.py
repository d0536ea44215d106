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
