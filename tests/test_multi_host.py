"""CPU: host logic of the sharded path under a real 2-process torch.distributed group (gloo)."""
import os
import subprocess
import sys
import numpy as np
import simsandgames_b200 as sg
from simsandgames_b200 import parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_partition_matches_library_rule():
    for ns, world in ((101, 2), (4096, 8), (16384, 8), (2050, 4), (7, 3)):
        part = parallel.slab_partition(ns, world)
        assert part[0][0] == 0 and part[-1][1] == ns
        assert all(a[1] == b[0] for a, b in zip(part, part[1:]))           # contiguous, no gaps, no overlap
        assert [sg.slab_rows(ns, r, world) for r in range(world)] == part  # same rule as the C side
        sizes = [b - a for a, b in part]
        assert max(sizes) - min(sizes) <= 1


def test_two_process_gloo_id_broadcast_and_gather(tmp_path):
    rng = np.random.default_rng(0)
    inp = str(tmp_path / "in.npz")
    a = rng.normal(size=(37, 12)).astype(np.float32)
    np.savez(inp, vel_fast=a)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29400 + os.getpid() % 500), os.path.join(ROOT, "tests", "multi_worker.py"), inp, str(tmp_path / "out"), "gloo"]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert np.array_equal(np.load(str(tmp_path / "out_stitched.npy")), a * 2.0)
