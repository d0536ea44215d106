"""Host-side logic of the N > 1 path on CPU: two gloo processes agree on the NCCL-id broadcast and on the interleaved partition that
tiles every wavefront step exactly once (no GPU, no compute calls)."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %r)
    import torch.distributed as dist
    from mvskit_b200 import dist as pdist, pmk
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    payload = bytes(range(128)) if rank == 0 else None
    got = pdist.broadcast_bytes(dist, payload, 128)
    assert got == bytes(range(128)), got[:8]
    gw, gh = 80, 60
    mine = []
    for d in range(gw + gh - 1):
        mine += pdist.step_tasks(gw, gh, d, rank, world)
    out = [None] * world
    dist.all_gather_object(out, mine)
    if rank == 0:
        cells = [c for part in out for c in part]
        assert len(cells) == len(set(cells)) == gw * gh, (len(cells), len(set(cells)))      # every dest cell exactly once
        assert abs(len(out[0]) - len(out[1])) <= gw + gh                                    # and evenly
        print("OK", world, len(cells))
    dist.barrier()
    dist.destroy_process_group()
""") % ROOT


def test_two_rank_partition_and_id_broadcast(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, PMK_FAST_BUILD=os.environ.get("PMK_FAST_BUILD", "0"))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "OK 2 4800" in r.stdout, r.stdout[-500:]
