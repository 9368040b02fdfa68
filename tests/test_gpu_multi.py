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


def _shard_case(tmp_path, make, world):
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
    return prob, want


@pytest.mark.parametrize("make", [flip_problem, euler_problem])
@pytest.mark.parametrize("world", [2, 4])
def test_multi_rank_shard_bit_identical_to_one_gpu(tmp_path, make, world):
    prob, want = _shard_case(tmp_path, make, world)
    parts = [np.load(str(tmp_path / f"out_rank{k}.npz")) for k in range(world)]
    for k in want:
        got = np.concatenate([pt[k] for pt in parts], axis=0)
        assert_bit_equal(got, want[k], f"{world}-rank shard {k}")
    # deep halo: one neighbour exchange every H-1 = 8 colour passes, each followed by ONE launch of the temporally blocked kernel
    assert all(int(pt["counters"][1]) == -(-2 * int(prob["iters"]) // 8) for pt in parts)
    assert all(int(pt["counters"][0]) <= 2 + -(-2 * int(prob["iters"]) // 8) for pt in parts)   # prep + one launch per exchange


def _launch(tmp_path, world, inp):
    port = 29600 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_worker.py"), inp, str(tmp_path / "out"), "gloo"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=150)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return [np.load(str(tmp_path / f"out_rank{k}.npz")) for k in range(world)]


@pytest.mark.parametrize("world,rows", [(2, None), (4, None), (2, [0, 33, 101])])
@pytest.mark.parametrize("start,steps,latch", [(0, 12, True), (40, 10, False)])
def test_sharded_flip_step_matches_one_gpu(tmp_path, world, rows, start, steps, latch):
    """The whole FLIP step, slab-sharded over `world` GPUs, against the one-GPU red-black run from the same state.
    With the rest density given (latch=False) every particle and every grid row is bit-identical; when the latch
    runs inside the sharded step the cross-rank sum may round differently, so that case is compared to 1e-5.
    rows: caller-chosen unequal slabs (flip_create_sharded_rows, the load-balancing entry point)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    o, p, g = co.make_flip_scene("default", oracle_kind="port")
    o.simulate(**p, n_steps=start)
    P = o.num_particles
    rest = 0.0 if latch else float(o.particle_rest_density)
    one = gpu_like(o, g)
    if latch:
        one.particle_rest_density = 0.0
    from tests.helpers import product_args
    one.simulate(**product_args(p), n_steps=steps, solver_order=sg.RED_BLACK, update_colors=False)
    inp = str(tmp_path / "in.npz")
    np.savez(inp, flip_steps=steps, pos=o.particle_pos[:2 * P], vel=o.particle_vel[:2 * P], s=o.s, u=o.u, v=o.v, rest=rest,
             capacity=2 * P, dt=p["dt"], iters=50, ox=p["ox"], oy=p["oy"], orad=p["orad"], **({"row_starts": np.array(rows)} if rows else {}),
             **{k: g[k] for k in ("density", "width", "height", "spacing", "radius")})
    parts = _launch(tmp_path, world, inp)
    ids = np.concatenate([pt["ids"] for pt in parts])
    assert np.array_equal(np.sort(ids), np.arange(P)), "every particle is owned by exactly one rank"
    pos = np.empty((P, 2), np.float32); vel = np.empty((P, 2), np.float32)
    pos[ids] = np.concatenate([pt["pos"] for pt in parts]).reshape(-1, 2)
    vel[ids] = np.concatenate([pt["vel"] for pt in parts]).reshape(-1, 2)
    want_pos, want_vel = one.readback("particle_pos").reshape(-1, 2), one.readback("particle_vel").reshape(-1, 2)
    ny, nx = o.f_num_y, o.f_num_x
    grids = {k: np.concatenate([pt[k] for pt in parts], axis=0) for k in ("u", "v", "p", "cell_type", "density")}
    want = {"u": one.readback("u"), "v": one.readback("v"), "p": one.readback("p"), "cell_type": one.readback("cell_type"),
            "density": one.readback("particle_density")}
    if latch:
        assert abs(float(parts[0]["rest"]) - float(one.particle_rest_density)) <= 1e-5 * float(one.particle_rest_density)
        assert np.abs(pos - want_pos).max() <= 1e-4 and np.abs(vel - want_vel).max() <= 5e-2
    else:
        assert_bit_equal(pos, want_pos, f"{world}-rank particle_pos")
        assert_bit_equal(vel, want_vel, f"{world}-rank particle_vel")
        for k in grids:
            assert_bit_equal(grids[k], want[k].reshape(ny, nx), f"{world}-rank {k}")
    assert all(int(pt["info"][4]) > 0 for pt in parts)   # NCCL exchanges happened


@pytest.mark.parametrize("start,steps", [(0, 6), (40, 6)])
def test_single_rank_sharded_flip_equals_plain_red_black(start, steps):
    """nranks=1 exercises the whole sharded code path (ownership filter, windows, row loops) without NCCL."""
    from tests.helpers import product_args
    o, p, g = co.make_flip_scene("default", oracle_kind="port")
    o.simulate(**p, n_steps=start)
    P = o.num_particles
    one = gpu_like(o, g)
    one.simulate(**product_args(p), n_steps=steps, solver_order=sg.RED_BLACK, update_colors=False)
    sh = sg.FlipFluidSharded(g["density"], g["width"], g["height"], g["spacing"], g["radius"], P, P, 0, 1)
    assert (sh.j0, sh.j1) == (0, o.f_num_y)
    for name in ("s", "u", "v"):
        sh.upload(name, getattr(o, name))
    sh.particle_rest_density = float(o.particle_rest_density)
    pos = o.particle_pos[:2 * P].reshape(-1, 2)
    assert sh.owns(pos).all()
    sh.upload_particles(pos, o.particle_vel[:2 * P], np.arange(P, dtype=np.int32))
    sh.simulate(**product_args(p), n_steps=steps)
    gp, gv, gi = sh.readback_particles()
    assert np.array_equal(np.sort(gi), np.arange(P))
    got = np.empty((P, 2), np.float32); got[gi] = gp.reshape(-1, 2)
    gotv = np.empty((P, 2), np.float32); gotv[gi] = gv.reshape(-1, 2)
    assert_bit_equal(got, one.readback("particle_pos"), "1-rank sharded pos")
    assert_bit_equal(gotv, one.readback("particle_vel"), "1-rank sharded vel")
    for name in ("u", "v", "p", "cell_type", "particle_density"):
        assert_bit_equal(sh.readback_own_rows(name), one.readback(name), f"1-rank sharded {name}")
