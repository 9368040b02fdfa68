"""GPU parity of the 3D Eulerian smoke step: simsandgames_b200.Fluid3DGPU -> libflipgpu.so against the CPU oracle
(oracle/_ref = the unmodified reference header eulerian_fluid/src/fluid3d.h where built, else the C port pinned to it).
Every call is bit-identical (-fmad=false, the reference's association order); red-black is bounded like the 2D solver."""
import numpy as np
import pytest

from oracle import cpu_oracle as co
import simsandgames_b200 as sg
from tests.helpers import assert_bit_equal

pytestmark = pytest.mark.gpu
FIELDS = ("u", "v", "w", "m", "pressure")


def gpu_like3(o):
    g = sg.Fluid3DGPU(o.num_x - 2, o.num_y - 2, o.num_z - 2, float(o.density), float(o.h))
    assert (g.num_x, g.num_y, g.num_z) == (o.num_x, o.num_y, o.num_z)
    for n in ("u", "v", "w", "m", "pressure", "solid"):
        g.upload(n, getattr(o, n))
    return g


@pytest.mark.parametrize("dims", [(30, 22, 26), (12, 9, 17), (20, 20, 20)])
def test_fluid3d_each_call_bit_identical(dims):
    o = co.make_fluid3d_scene(*dims, oracle_kind=co.best_kind("fluid3d"))
    g = gpu_like3(o)
    dt = 1 / 60.0
    for fn_o, fn_g, args in (("solve", "solveIncompressibility", (15, dt)), ("extrapolate", "extrapolate", ()),
                             ("advect_vel", "advectVel", (dt,)), ("advect_smoke", "advectSmoke", (dt,)), ("solve", "solveIncompressibility", (1, dt))):
        getattr(o, fn_o)(*args)
        getattr(g, fn_g)(*args)
        for n in FIELDS:
            assert_bit_equal(g.readback(n), getattr(o, n), f"{fn_o} {n} {dims}")


def test_fluid3d_steps_and_sample_field():
    o = co.make_fluid3d_scene(oracle_kind=co.best_kind("fluid3d"))
    g = gpu_like3(o)
    o.step(20, 1 / 60.0, n_steps=5)
    g.step(20, 1 / 60.0, n_steps=5, solver_order=sg.LEX_WAVEFRONT)
    for n in FIELDS:
        assert_bit_equal(g.readback(n), getattr(o, n), f"5 steps {n}")
    rng = np.random.default_rng(3)
    L = float(o.h) * np.array([o.num_x, o.num_y, o.num_z], np.float32)
    q = (rng.uniform(-0.1, 1.1, (200, 3)) * L).astype(np.float32)
    for fld in range(4):
        got = g.sampleField(q[:, 0], q[:, 1], q[:, 2], fld)
        want = np.array([o.sample(float(a), float(b), float(c), fld) for a, b, c in q], np.float32)
        assert_bit_equal(got, want, f"sampleField field {fld}")


def test_fluid3d_red_black_shares_the_fixed_point():
    """T2 in 3D: red-black and the reference order converge to the same projection (distance at the fp32 floor)."""
    o = co.make_fluid3d_scene(oracle_kind=co.best_kind("fluid3d"))
    g = gpu_like3(o)
    o.solve(1500, 1 / 60.0)
    g.solveIncompressibility(1500, 1 / 60.0, solver_order=sg.RED_BLACK)
    scale = max(float(np.abs(o.u).max()), 1e-3)
    for n in ("u", "v", "w"):
        assert float(np.abs(g.readback(n) - getattr(o, n)).max()) <= 5e-4 * scale, n
