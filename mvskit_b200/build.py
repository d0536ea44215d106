"""Build mvskit_b200/libpmk.so (the C-ABI shared library) for sm_100a with nvcc.

In-tree build: the .so is git-ignored but travels with the repository snapshot to the GPU box.
nvcc cross-compiles without a GPU, so this also runs in the CPU-only build container.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libpmk.so")
SOURCES = ["pmk_api.cu"]
HEADERS = ["pmk_device.cuh", "pmk_ncc.cuh", os.path.join("..", "..", "include", "pmk.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(CSRC, "pmk_exports.map"), __file__]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


FAST_MARK = OUT + ".fast"


def build(force: bool = False, verbose: bool = False, fast: bool = False) -> str:
    """fast=True: development build with the wsize-7 kernels only (4x quicker to compile); a later plain build() replaces it."""
    fast = fast or os.environ.get("PMK_FAST_BUILD") == "1"
    if os.path.exists(FAST_MARK) and not fast:
        force = True
    if not force and not stale():
        return OUT
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC,-ffp-contract=off", "-Xptxas", "-v" if verbose else "-warn-spills",
           "-shared", "-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES] + \
          ["-Xlinker", "--version-script=" + os.path.join(CSRC, "pmk_exports.map")]
    if fast:
        cmd.insert(1, "-DPMK_WS_ONLY=7")
    for flag in os.environ.get("PMK_NVCC_FLAGS", "").split():
        cmd.insert(1, flag)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libpmk.so")
    if fast:
        open(FAST_MARK, "w").close()
    elif os.path.exists(FAST_MARK):
        os.remove(FAST_MARK)
    return OUT


HOST_DIR = os.path.join(HERE, "host")
HOST_EXE = os.path.join(HOST_DIR, "pmmvps_b200")


def build_host(force: bool = False) -> str:
    """g++ build of the host-side mirror of the reference's classes + driver (mvskit_b200/host/), linked against libpmk.so."""
    srcs = [os.path.join(HOST_DIR, f) for f in ("pmmvps.cpp", "main.cpp")]
    deps = srcs + [os.path.join(HOST_DIR, "pmmvps.hpp"), os.path.join(HERE, "..", "include", "pmk.h")]
    if not force and os.path.exists(HOST_EXE) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_EXE) for d in deps):
        return HOST_EXE
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    cmd = [cxx, "-std=c++14", "-O2", "-ffp-contract=off", "-Wall", "-o", HOST_EXE] + srcs + ["-L" + HERE, "-lpmk", "-Wl,-rpath,$ORIGIN/.."]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building the host driver")
    return HOST_EXE


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv, fast="--fast" in sys.argv))
    print(build_host(force=True))
