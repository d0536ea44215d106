"""Shared fixtures.  `-m "not gpu"` runs here on CPU; `-m gpu` runs on a B200 through the C ABI."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu() -> bool:
    try:
        import ctypes
        lib = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int()
        return lib.cuInit(0) == 0 and lib.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def small_scene():
    """config 1 (textured plane, 5 views) at half size: 320x240."""
    from mvskit_b200 import synth
    return synth.make_scene(1, scale=0.5).render()


@pytest.fixture(scope="session")
def sphere_scene():
    """config 2 shape (sphere on floor, occlusions) with 9 views at half size."""
    from mvskit_b200 import synth
    return synth.make_scene(2, scale=0.5, nviews=9).render()


@pytest.fixture(scope="session")
def scene_dir(small_scene, tmp_path_factory):
    from mvskit_b200 import synth
    d = tmp_path_factory.mktemp("scene_c1")
    return synth.write_scene(small_scene, str(d))


@pytest.fixture(scope="session")
def coracle(small_scene):
    from oracle import pyoracle
    pyoracle.build(ref=False)
    return pyoracle.COracle(small_scene.P, small_scene.images)


@pytest.fixture(scope="session")
def coracle_sphere(sphere_scene):
    from oracle import pyoracle
    pyoracle.build(ref=False)
    return pyoracle.COracle(sphere_scene.P, sphere_scene.images)


@pytest.fixture(scope="session")
def reflib(scene_dir):
    """The reference's own code (oracle/_ref/libpmref.so).  One scene per process."""
    from oracle import pyoracle
    if not os.path.exists(pyoracle.REF_SO):
        if os.path.isdir("/root/reference/pmmvps"):
            pyoracle.build(ref=True)
        else:
            pytest.skip("oracle/_ref/libpmref.so not built and /root/reference absent")
    return pyoracle.RefLib(scene_dir)


@pytest.fixture(scope="session")
def hyps(small_scene):
    """Half well-observed hypotheses (all tau views valid), half unfiltered draws (missing / back-facing views)."""
    a = small_scene.hypotheses(2048, seed=7, well_observed=True)
    b = small_scene.hypotheses(2048, seed=8, well_observed=False, normal_jitter_deg=35.0)
    return tuple(np.concatenate([x, y]) for x, y in zip(a, b))


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def assert_bits_equal(a, b, what=""):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.dtype.kind == "f":
        bad = np.nonzero(bits(a) != bits(b))
    else:
        bad = np.nonzero(a != b)
    assert bad[0].size == 0, f"{what}: {bad[0].size} of {a.size} differ, first at {tuple(x[0] for x in bad)}: {a[tuple(x[0] for x in bad)]} vs {b[tuple(x[0] for x in bad)]}"
