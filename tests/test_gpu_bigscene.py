"""GPU parity at a benchmark-scale grid: the synthetic dam-break rule of BASELINE configs[2]/[4] at 1024^2 cells /
3 990 320 particles WITH the disc obstacle, and the Eulerian wind tunnel at 1026^2 cells, against the CPU oracle
(oracle/_ref = the unmodified reference where it was built, else the C port pinned to it) on identical inputs.

The default-tank tests (test_gpu_flip.py) run on 134x101 cells: one solver tile column, a handful of particles per warp.
Here every kernel runs many waves of tiles / CTAs, all inter-tile dependencies of the blocked wavefront are active, the
obstacle cuts through tiles, and the particle state is a disordered splash (seeded velocity noise on top of three
reference steps), so crowded hash cells and multi-cell moves occur.  Same tiers as SURVEY 8c: T0 bit-exact bins,
T1 same-state replay, T4 short trajectory.  One reference step at this size takes ~2 s on the host (50 sweeps).
"""
import numpy as np
import pytest

from oracle import cpu_oracle as co
import simsandgames_b200 as sg
from tests.helpers import gpu_like, clone_oracle, assert_bit_equal, product_args

pytestmark = pytest.mark.gpu

N = 1024


@pytest.fixture(scope="module")
def big():
    kind = co.best_kind("flip")
    o, p, g = co.make_flip_scene("synthetic", N=N, oracle_kind=kind, with_obstacle=True)
    assert (o.f_num_x, o.f_num_y, o.num_particles) == (N, N, 3990320)
    p50 = dict(p, num_pressure_iters=50)
    P = o.num_particles
    rng = np.random.default_rng(20261020)
    o.particle_vel[:2 * P] = rng.normal(0.0, 0.6, 2 * P).astype(np.float32)   # ~1.7 radii of random motion per step
    o.simulate(**p50, n_steps=3)
    return o, p50, g, kind


def test_big_integrate_bin_separate_collide(big):
    """F1 + F2 bit-exact (T0), F3 kernel == CPU model of its schedule, F4 bit-exact, at 4 M particles."""
    base, p, g, kind = big
    o = clone_oracle(base, g, kind="port")
    sim = gpu_like(o, g)
    P = o.num_particles
    o.integrate(p["dt"], p["gravity"]); sim.integrateParticles(p["dt"], p["gravity"])
    assert_bit_equal(sim.readback("particle_pos"), o.particle_pos[:2 * P], "F1 pos")
    assert_bit_equal(sim.readback("particle_vel"), o.particle_vel[:2 * P], "F1 vel")
    o.bin_particles(); sim.binParticles()
    assert int(o.num_cell_particles.max()) >= 3          # the splash state has crowded hash cells
    assert_bit_equal(sim.readback("num_cell_particles"), o.num_cell_particles, "F2 num_cell_particles")
    assert_bit_equal(sim.readback("first_cell_particle"), o.first_cell_particle, "F2 first_cell_particle")
    assert_bit_equal(sim.readback("cell_particle_ids"), o.cell_particle_ids[:P], "F2 cell_particle_ids")
    o.separate_jacobi(2); sim.separateParticles(2)
    assert_bit_equal(sim.readback("particle_pos"), o.particle_pos[:2 * P], "F3 pos vs the schedule's CPU model")
    o.collide(p["ox"], p["oy"], 0.25, -0.5, p["orad"])
    sim.handleParticleCollisions(p["ox"], p["oy"], 0.25, -0.5, p["orad"])
    assert_bit_equal(sim.readback("particle_pos"), o.particle_pos[:2 * P], "F4 pos")
    assert_bit_equal(sim.readback("particle_vel"), o.particle_vel[:2 * P], "F4 vel")


@pytest.fixture(scope="module")
def big_pre_solve(big):
    """reference state right before solveIncompressibility of the 4th step"""
    base, p, g, kind = big
    o = clone_oracle(base, g, kind="port")
    o.integrate(p["dt"], p["gravity"]); o.push_apart(2); o.collide(p["ox"], p["oy"], 0.0, 0.0, p["orad"])
    pre_p2g = clone_oracle(o, g, kind="port")
    o.transfer(True); o.update_density()
    return pre_p2g, o


def test_big_p2g_density_cell_types(big, big_pre_solve):
    """F5 bit-exact, F6/F7 <= 1e-5 relative (order of the adds into a node), incl. the cells around the obstacle."""
    _, p, g, _ = big
    pre, post = big_pre_solve
    sim = gpu_like(pre, g)
    sim.transferVelocitiesToGrid()
    assert_bit_equal(sim.readback("cell_type"), post.cell_type, "F5 cell_type")
    for name in ("u", "v", "particle_density"):
        a, b = sim.readback(name), getattr(post, name)
        scale = max(1.0, float(np.abs(b).max()))
        assert np.abs(a - b).max() <= 1e-5 * scale, (name, float(np.abs(a - b).max()), scale)
    assert abs(float(sim.particle_rest_density) - float(post.particle_rest_density)) <= 1e-5 * float(post.particle_rest_density)


@pytest.mark.parametrize("iters", [50, 200])
def test_big_solver_lexicographic_bit_identical(big, big_pre_solve, iters):
    """F9 at the reference's 50 and 200 sweeps (configs[2] / configs[4]) with the obstacle: u, v, p bit-identical to
    flip_fluid.h:451-495 -- 32 x 32 tiles (or strips) per chunk, every inter-tile dependency exercised."""
    _, p, g, _ = big
    _, pre = big_pre_solve
    o = clone_oracle(pre, g, kind="port")
    sim = gpu_like(o, g)
    o.solve(iters, p["dt"], 1.9, True)
    sim.solveIncompressibility(iters, p["dt"], 1.9, True, solver_order=sg.LEX_WAVEFRONT)
    for name in ("u", "v", "p", "prev_u", "prev_v"):
        assert_bit_equal(sim.readback(name), getattr(o, name), f"F9 {name} at {iters} sweeps")
    if iters == 50:   # F8 on the solved field: gather, same arithmetic -> bit-exact
        o.transfer(False, 0.9)
        sim.transferVelocitiesToParticles(0.9)
        P = o.num_particles
        assert_bit_equal(sim.readback("particle_vel"), o.particle_vel[:2 * P], "F8 vel")


def test_big_short_trajectory_reference_order(big):
    """T4: two whole steps, reference solver order, separation off -> only the order of the P2G adds differs."""
    base, p, g, _ = big
    o = clone_oracle(base, g, kind="port")
    sim = gpu_like(o, g)
    pp = dict(p, separate_particles=False)
    o.simulate(**pp, n_steps=2)
    sim.simulate(**product_args(pp), n_steps=2, solver_order=sg.LEX_WAVEFRONT, update_colors=False)
    P = o.num_particles
    dpos = float(np.abs(sim.readback("particle_pos") - o.particle_pos[:2 * P]).max())
    assert dpos <= 4e-5, dpos                      # domain width 3
    assert_bit_equal(sim.readback("cell_type"), o.cell_type, "cell_type after 2 steps")


def test_big_eulerian_steps_bit_identical():
    """E1-E5 at 1026^2 cells: two caller frames (project 40 -> extrapolate -> advectVel -> advectSmoke, fluid_ui.h:107-111)
    bit-identical in u, v, m, pressure."""
    kind = co.best_kind("euler")
    e = co.make_euler_scene(N, N, oracle_kind=kind)
    es = sg.scenes.make_wind_tunnel(N, N)
    assert (es.getNumX(), es.getNumY()) == (N + 2, N + 2)
    e.step(40, 1 / 60.0, n_steps=2)
    es.step(40, 1 / 60.0, n_steps=2, solver_order=sg.LEX_WAVEFRONT)
    for name in ("u", "v", "m", "pressure"):
        assert_bit_equal(es.readback(name), getattr(e, name), f"Eulerian {name}")
