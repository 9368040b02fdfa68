"""GPU: the slab-sharded red-black projection.  Single-rank cases run on any B200 box; the 2-rank
NCCL case needs two GPUs (gpurun --gpus 2) and is skipped otherwise."""
import os
import subprocess
import sys
import numpy as np
import pytest
import torch

from oracle import cpu_oracle as co
import simsandgames_b200 as sg
from tests.helpers import gpu_like, clone_oracle, assert_bit_equal

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def flip_problem(step=40):
    o, p, g = co.make_flip_scene("default", oracle_kind="port")
    o.simulate(**p, n_steps=step)
    o.integrate(p["dt"], p["gravity"]); o.push_apart(2); o.collide(p["ox"], p["oy"], 0, 0, p["orad"])
    o.transfer(True); o.update_density()
    ny, nx = o.f_num_y, o.f_num_x
    cp = float(np.float32(o.density) * np.float32(o.h) / np.float32(p["dt"]))
    prob = dict(vel_fast=o.u.reshape(ny, nx).copy(), vel_slow=o.v.reshape(ny, nx).copy(), p=np.zeros((ny, nx), np.float32),
                s=o.s.reshape(ny, nx).copy(), density=o.particle_density.reshape(ny, nx).copy(),
                cell_type=o.cell_type.reshape(ny, nx).copy(), x_fast=True, iters=50, cp=cp, omega=1.9, drift=True,
                rest=float(o.particle_rest_density))
    sim = gpu_like(o, g)
    sim.solveIncompressibility(50, p["dt"], 1.9, True, solver_order=sg.RED_BLACK)
    want = {k: sim.readback(n).reshape(ny, nx) for k, n in (("vel_fast", "u"), ("vel_slow", "v"), ("p", "p"))}
    return prob, want


def euler_problem():
    e = co.make_euler_scene(96, 72, oracle_kind="port")
    e.step(40, 1 / 60.0, n_steps=20)
    sim = sg.FluidGPU(96, 72, 1000.0, 0.01)
    for n in ("u", "v", "m", "solid"):
        sim.upload(n, getattr(e, n))
    sim.solveIncompressibility(40, 1 / 60.0, solver_order=sg.RED_BLACK)
    nx, ny = e.num_x, e.num_y
    cp = float(np.float32(1000.0) * np.float32(0.01) / np.float32(1 / 60.0))
    prob = dict(vel_fast=e.v.copy(), vel_slow=e.u.copy(), p=np.zeros((nx, ny), np.float32), s=(1 - e.solid).astype(np.float32),
                density=np.zeros((nx, ny), np.float32), cell_type=np.where(e.solid != 0, 2, 0).astype(np.int32), x_fast=False,
                iters=40, cp=cp, omega=1.9, drift=False, rest=0.0)
    want = dict(vel_fast=sim.readback("v"), vel_slow=sim.readback("u"), p=sim.readback("pressure"))
    return prob, want


@pytest.mark.parametrize("make", [flip_problem, euler_problem])
def test_single_rank_shard_equals_red_black(make):
    prob, want = make()
    ns, nf = prob["vel_fast"].shape
    sh = sg.ShardedProjection(nf, ns, prob["x_fast"], 0, 1)
    assert (sh.j0, sh.j1) == (0, ns)
    for name in ("vel_fast", "vel_slow", "p", "s", "density", "cell_type"):
        sh.upload_window(name, prob[name])
    sh.solve(prob["iters"], prob["cp"], prob["omega"], prob["drift"], prob["rest"])
    for k in want:
        assert_bit_equal(sh.readback_own(k), want[k], f"1-rank shard {k}")


@pytest.mark.parametrize("make", [flip_problem, euler_problem])
@pytest.mark.parametrize("world", [2, 4])
def test_multi_rank_shard_bit_identical_to_one_gpu(tmp_path, make, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    prob, want = make()
    inp = str(tmp_path / "in.npz")
    np.savez(inp, **prob)
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_worker.py"), inp, str(tmp_path / "out"), "gloo"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    parts = [np.load(str(tmp_path / f"out_rank{k}.npz")) for k in range(world)]
    for k in want:
        got = np.concatenate([pt[k] for pt in parts], axis=0)
        assert_bit_equal(got, want[k], f"{world}-rank shard {k}")
    # deep halo: one neighbour exchange every H-1 = 8 colour passes
    assert all(int(pt["counters"][1]) == -(-2 * int(prob["iters"]) // 8) for pt in parts)
