"""The host-side mirror of the reference's classes (mvskit_b200/host: Option / PmMvps / PatchManager / Patch) driving the whole
path through the C ABI: the reference's own driver (test/test.cpp) on the reference's on-disk layout."""
import os
import subprocess

import numpy as np
import pytest

from conftest import HAVE_GPU

# the compiled reference on the same scene and seeds, iteration 0 (tools/ref_pipeline.py): 27321 patches, filter removes none,
# |z| / scene scale quantiles 50 / 90 % = 5.98e-4 / 1.29e-3
REF_ITER0_PATCHES = 27321
REF_ITER0_Q = (5.98e-4, 1.29e-3)


def _exe():
    from mvskit_b200 import build
    build.build()
    return build.build_host()


def read_patch_file(path):
    tok = open(path).read().split()
    assert tok[0] == "PATCHES"
    n = int(tok[1])
    i = 2
    coord, ncc, nimg = np.zeros((n, 4), np.float32), np.zeros(n, np.float32), np.zeros(n, np.int32)
    for p in range(n):
        assert tok[i] == "PATCHES"
        coord[p] = [float(t) for t in tok[i + 1:i + 5]]
        ncc[p] = float(tok[i + 9])
        k = int(tok[i + 12])
        nimg[p] = k
        i += 13 + k
        i += 1 + int(tok[i])
    return coord, ncc, nimg


def test_driver_fails_loudly_without_a_gpu(scene_dir):
    if HAVE_GPU:
        pytest.skip("a GPU is present")
    r = subprocess.run([_exe(), scene_dir], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr, (r.returncode, r.stderr[-300:])


def test_option_errors_match_the_reference(tmp_path):
    """Unknown keys and a missing `images` line are fatal (cerr + exit(1)), option.cpp:121-133."""
    d = tmp_path / "bad"
    d.mkdir()
    (d / "option").write_text("image 5\nbogus 1\n")
    r = subprocess.run([_exe(), str(d) + "/"], capture_output=True, text=True)
    assert r.returncode == 1 and "Unrecognizable option: bogus" in r.stderr
    (d / "option").write_text("image 5\nlevel 1\n")
    r = subprocess.run([_exe(), str(d) + "/"], capture_output=True, text=True)
    assert r.returncode == 1 and "m_flag not specified" in r.stderr


@pytest.mark.gpu
def test_pmmvps_run_end_to_end(scene_dir, small_scene):
    r = subprocess.run([_exe(), scene_dir, "--group", "1"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    n = int(r.stdout.split()[-1])
    ply = os.path.join(scene_dir, "ply")
    for it in range(3):
        assert os.path.exists(os.path.join(ply, f"refined_patches_before_refine_{it}.ply"))
        assert os.path.exists(os.path.join(ply, f"refined_patches_{it}.ply"))
    head = open(os.path.join(ply, "refined_patches_0.ply")).read().split("\n")
    n0 = int(head[2].split()[-1])
    assert head[0] == "ply" and head[2].startswith("element vertex")
    # iteration 0 against the reference's own run from the same seeds
    assert abs(n0 - REF_ITER0_PATCHES) <= 0.05 * REF_ITER0_PATCHES, n0
    pts = np.array([[float(t) for t in ln.split()[:3]] for ln in head[13:13 + n0]])
    q = np.quantile(np.abs(pts[:, 2]) / small_scene.scene_scale, [0.5, 0.9])
    assert q[0] <= 1.15 * REF_ITER0_Q[0] and q[1] <= 1.15 * REF_ITER0_Q[1], q
    # final state
    coord, ncc, nimg = read_patch_file(os.path.join(ply, "final.patch"))
    assert len(coord) == n and n > 0.9 * n0
    assert (nimg >= 3).all() and np.median(ncc) > 0.95
    z = np.abs(coord[:, 2]) / small_scene.scene_scale
    assert np.quantile(z, 0.9) <= 2.0e-3, np.quantile(z, [0.5, 0.9, 0.99])


@pytest.mark.gpu
def test_patch_manager_pass_throughs_of_the_host_mirror(scene_dir):
    """`pmmvps_b200 <prefix> --selftest`: the C++ mirror's PatchManager methods on the loaded seeds -- one-patch computeNcc (wide call) against the
    batched computeNcc (Patch objects marshalled into the byte-lean wire format) bit for bit, sortPatches, setScales, isVisible0, findNeighbors,
    removePatch / updateDepthMaps by collect index on a store that is not in collect order, syncGrids."""
    r = subprocess.run([_exe(), scene_dir, "--selftest"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "selftest ok" in r.stdout, (r.stdout[-500:], r.stderr[-1500:])


def test_patch_file_round_trip_matches_what_the_reference_reads(scene_dir, reflib, tmp_path):
    """The mirror's Patch stream operators (patch.cpp:31-88 format) on CPU: parse the seed file, write it back, and compare the
    records with what the reference's own readPatches got from the same file."""
    src = os.path.join(scene_dir, "ply", "00000000.patch")
    out = str(tmp_path / "roundtrip.patch")
    r = subprocess.run([_exe(), "--patch-io", src, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    coord, ncc, nimg = read_patch_file(out)
    reflib.clear_patches()
    reflib.set_depth(0)
    reflib.create_patches()
    rb = reflib.get_patches()
    assert len(coord) == rb.n
    # the reference's m_ppatches is in collect order, the file in write order: compare as sets of (coord, images count)
    # (operator<< prints 6 significant digits, so the written coordinates agree to ~1e-5 of their magnitude)
    ia, ib = np.lexsort(np.round(coord[:, :3], 3).T[::-1]), np.lexsort(np.round(rb.coord[:, :3], 3).T[::-1])
    assert np.abs(coord[ia, :3] - rb.coord[ib, :3]).max() <= 1e-4
    assert np.array_equal(nimg[ia], rb.nimages[ib])
    # and a second pass through the operators is the identity on the text
    out2 = str(tmp_path / "roundtrip2.patch")
    assert subprocess.run([_exe(), "--patch-io", out, out2]).returncode == 0
    assert open(out).read() == open(out2).read()
