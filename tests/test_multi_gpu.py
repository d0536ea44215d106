"""N > 1 on real GPUs (skipped on a single-GPU box): two ranks, every step's dest cells dealt out to the ranks in turn, NCCL all-gather of the step mutations.
Every replica must end with the same store, and that store must be the single-GPU store bit for bit."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count() -> int:
    try:
        import ctypes
        lib = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int()
        return n.value if lib.cuInit(0) == 0 and lib.cuDeviceGetCount(ctypes.byref(n)) == 0 else 0
    except OSError:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs")
def test_two_rank_pipeline_equals_single_gpu():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tools", "run_pipeline_mg.py"), "--config", "1", "--scale", "0.5",
                        "--iters", "2"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert out["n_gpus"] == 2 and out["replicas_identical"] and out["equals_single_gpu"], out
    assert out["patches"] == out["single_gpu_patches"] > 20000
