"""Scene construction for the headless drivers (what the reference UI shells do in OnUserCreate).

default tank  : flip_fluid/src/main.cpp:33-79 (640x480 window, res 100, obstacle (3,2) r=.15)  -> config S0
synthetic tank: SURVEY 8d rule for the large configs: sp = 3/N, W = H = (N-.5)*sp, r = .19*sp,
                hex block rel_w=.6, rel_h=.8 (main.cpp:43-66), dt = (1/60)*(100/N); optional static disc
                obstacle at (.75W,.67H) radius .05W with 200 pressure iterations (config S4).
wind tunnel   : eulerian_fluid/src/fluid_ui.h:64-89 scaled to an x-by-y interior.
The particle lattice and tank walls come from the C-ABI helper flip_scene_tank (host C++ fp32, the
reference's own expressions); the few scalars here are numpy fp32.
"""
import ctypes as C
import numpy as np
from .binding import FlipFluidGPU, FluidGPU, load_library, _check

f32 = np.float32


def tank_geometry(kind="default", N=None, with_obstacle=None):
    if kind == "default":
        sim_h = f32(3)
        c_scale = f32(480) / sim_h
        W, H = f32(640) / c_scale, sim_h
        sp = H / f32(100)
        r = f32(.3) * sp
        dt = f32(1) / f32(60)
        obstacle = (f32(3), f32(2), f32(.15))
        iters = 50
    elif kind == "synthetic":
        sp = f32(3) / f32(N)
        W = H = (f32(N) - f32(.5)) * sp
        r = f32(.19) * sp
        dt = (f32(1) / f32(60)) * (f32(100) / f32(N))
        obstacle = (f32(.75) * W, f32(.67) * H, f32(.05) * W) if with_obstacle else None
        iters = 200 if with_obstacle else 50
    else:
        raise ValueError(kind)
    return dict(width=float(W), height=float(H), spacing=float(sp), radius=float(r), dt=float(dt),
                obstacle=None if obstacle is None else tuple(float(t) for t in obstacle), iters=iters)


def tank_particles(width, height, spacing, radius, f_num_x=0, f_num_y=0, rel_w=0.6, rel_h=0.8, want_s=True, pos_out=None):
    """(pos float32[2P], s float32[ny*nx] or None, P) via flip_scene_tank."""
    lib = load_library()
    n = C.c_int()
    args = [C.c_float(width), C.c_float(height), C.c_float(spacing), C.c_float(radius), C.c_float(rel_w), C.c_float(rel_h),
            C.c_int(f_num_x), C.c_int(f_num_y)]
    _check(lib.flip_scene_tank(*args, None, C.c_size_t(0), None, C.byref(n)), "flip_scene_tank")
    P = n.value
    pos = pos_out if pos_out is not None else np.empty(2 * P, np.float32)
    s = np.empty(f_num_x * f_num_y, np.float32) if (want_s and f_num_x > 0) else None
    _check(lib.flip_scene_tank(*args, pos.ctypes.data_as(C.c_void_p), C.c_size_t(pos.size),
                               s.ctypes.data_as(C.c_void_p) if s is not None else None, C.byref(n)), "flip_scene_tank")
    return pos, s, P


def make_flip_tank(kind="default", N=None, with_obstacle=None, device=-1):
    """Returns (FlipFluidGPU, simulate-kwargs, geometry dict) initialised like OnUserCreate."""
    g = tank_geometry(kind, N, with_obstacle)
    P = _count_only(g)
    sim = FlipFluidGPU(1000.0, g["width"], g["height"], g["spacing"], g["radius"], P, device=device)
    pos, s, P2 = tank_particles(g["width"], g["height"], g["spacing"], g["radius"], sim.f_num_x, sim.f_num_y)
    assert P2 == P
    sim.num_particles = P
    sim.upload("particle_pos", pos)
    sim.upload("s", s)
    ox = oy = orad = 0.0
    if g["obstacle"] is not None:
        ox, oy, orad = g["obstacle"]
        sim.set_obstacle(ox, oy, 0.0, 0.0, orad)  # setObstacle(x, y, reset=true), main.cpp:79
    params = dict(dt=g["dt"], gravity=9.81, flip_ratio=0.9, num_pressure_iters=g["iters"], num_particle_iters=2,
                  over_relaxation=1.9, compensate_drift=True, separate_particles=True,
                  obstacle_x=ox, obstacle_y=oy, obstacle_vel_x=0.0, obstacle_vel_y=0.0, obstacle_radius=orad)
    g = dict(g, density=1000.0, max_particles=P)
    return sim, params, g


def _count_only(g):
    lib = load_library()
    n = C.c_int()
    _check(lib.flip_scene_tank(C.c_float(g["width"]), C.c_float(g["height"]), C.c_float(g["spacing"]), C.c_float(g["radius"]),
                               C.c_float(0.6), C.c_float(0.8), C.c_int(0), C.c_int(0), None, C.c_size_t(0), None, C.byref(n)),
           "flip_scene_tank")
    return n.value


def wind_tunnel_arrays(num_x, num_y):
    """Initial u, m, solid of eulerian_fluid/src/fluid_ui.h:64-89 (setObstacle :25-62 with reset=true)."""
    u = np.zeros((num_x, num_y), np.float32)
    m = np.ones((num_x, num_y), np.float32)
    solid = np.zeros((num_x, num_y), np.uint8)
    u[1, :] = 2.0
    solid[0, :] = 1
    solid[:, 0] = 1
    solid[:, -1] = 1
    ox, oy, r = num_x // 3, num_y // 2, num_x // 8
    i = np.arange(num_x)[:, None]
    j = np.arange(num_y)[None, :]
    disc = (i - ox) ** 2 + (j - oy) ** 2 < r * r
    disc[0, :] = False; disc[-1, :] = False; disc[:, 0] = False; disc[:, -1] = False
    solid[disc] = 1
    m[disc] = 1.0
    u[disc] = 0.0
    pipe_h = num_y // 6
    lo, hi = num_y // 2 - pipe_h // 2, num_y // 2 + pipe_h // 2
    m[0, lo:hi + 1] = 0.0
    return u, m, solid


def make_wind_tunnel(x=96, y=72, device=-1):
    sim = FluidGPU(x, y, 1000.0, 1 / 100.0, device=device)
    u, m, solid = wind_tunnel_arrays(sim.getNumX(), sim.getNumY())
    sim.upload("u", u)
    sim.upload("m", m)
    sim.upload("solid", solid)
    return sim
