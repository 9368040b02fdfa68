"""TEST INFRASTRUCTURE ONLY -- ctypes front-end for the CPU checkers in oracle/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (simsandgames_b200) never does.

Two interchangeable back-ends with the same phase-by-phase API:
  kind="ref"   oracle/_ref/libref_{flip,euler}.so -- the UNMODIFIED reference headers
               (/root/reference/flip_fluid/src/flip_fluid.h, eulerian_fluid/src/fluid.h)
               compiled by oracle/Makefile behind ref_harness_*.cpp
  kind="port"  oracle/_build/liboracle.so -- the plain-C restatement
               (flip_oracle.c, euler_oracle.c), pinned bit-exactly to "ref"
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATHS = {
    ("flip", "ref"): os.path.join(_HERE, "_ref", "libref_flip.so"),
    ("euler", "ref"): os.path.join(_HERE, "_ref", "libref_euler.so"),
    ("fluid3d", "ref"): os.path.join(_HERE, "_ref", "libref_fluid3d.so"),
    ("fluid3d", "port"): os.path.join(_HERE, "_build", "liboracle.so"),
    ("flip", "port"): os.path.join(_HERE, "_build", "liboracle.so"),
    ("euler", "port"): os.path.join(_HERE, "_build", "liboracle.so"),
}
_PREFIX = {("flip", "ref"): "refflip_", ("flip", "port"): "orcflip_",
           ("euler", "ref"): "refeuler_", ("euler", "port"): "orceuler_",
           ("fluid3d", "ref"): "reffluid3d_", ("fluid3d", "port"): "orcfluid3d_"}


def build(quiet=True):
    """make -C oracle (port always; _ref only when /root/reference exists)."""
    subprocess.run(["make", "-C", _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def available(family, kind):
    return os.path.exists(_PATHS[(family, kind)])


def best_kind(family="flip"):
    return "ref" if available(family, "ref") else "port"


_f32 = np.float32
_FLIP_FIELDS = {  # name -> (dtype, size expression)
    "u": ("f4", "C"), "v": ("f4", "C"), "du": ("f4", "C"), "dv": ("f4", "C"),
    "prev_u": ("f4", "C"), "prev_v": ("f4", "C"), "p": ("f4", "C"), "s": ("f4", "C"),
    "cell_type": ("i4", "C"), "cell_color": ("f4", "3C"),
    "particle_pos": ("f4", "2M"), "particle_vel": ("f4", "2M"), "particle_color": ("f4", "3M"),
    "particle_density": ("f4", "C"),
    "num_cell_particles": ("i4", "H"), "first_cell_particle": ("i4", "H1"), "cell_particle_ids": ("i4", "M"),
}


class FlipOracle:
    """Phase-by-phase CPU FLIP step (flip_fluid.h:156-581) with numpy views of the state."""

    def __init__(self, density, width, height, spacing, radius, max_particles, kind=None):
        kind = kind or best_kind("flip")
        self.kind = kind
        self.lib = C.CDLL(_PATHS[("flip", kind)])
        self._p = _PREFIX[("flip", kind)]
        fn = self._fn("create")
        fn.restype = C.c_void_p
        fn.argtypes = [C.c_float] * 5 + [C.c_int]
        self.h_ = C.c_void_p(fn(density, width, height, spacing, radius, max_particles))
        self._refresh_dims()
        sizes = {"C": self.f_num_cells, "3C": 3 * self.f_num_cells, "2M": 2 * max_particles,
                 "3M": 3 * max_particles, "M": max_particles, "H": self.p_num_cells, "H1": self.p_num_cells + 1}
        ptr = self._fn("ptr")
        ptr.restype = C.c_void_p
        ptr.argtypes = [C.c_void_p, C.c_char_p]
        for name, (dt, sz) in _FLIP_FIELDS.items():
            addr = ptr(self.h_, name.encode())
            n = sizes[sz]
            ctype = C.c_float if dt == "f4" else C.c_int
            arr = np.ctypeslib.as_array(C.cast(addr, C.POINTER(ctype)), shape=(max(n, 1),))[:n]
            setattr(self, name, arr)

    def _fn(self, name):
        return getattr(self.lib, self._p + name)

    def _refresh_dims(self):
        ints = (C.c_int * 8)()
        fl = (C.c_float * 8)()
        fn = self._fn("dims")
        fn.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)]
        fn(self.h_, ints, fl)
        (self.f_num_x, self.f_num_y, self.f_num_cells, self.p_num_x, self.p_num_y, self.p_num_cells,
         self.max_particles, self.num_particles) = list(ints)
        self.density, self.h, self.f_inv_spacing, self.particle_radius, self.p_inv_spacing, \
            self.particle_rest_density = [_f32(x) for x in list(fl)[:6]]

    def close(self):
        if self.h_:
            fn = self._fn("destroy")
            fn.argtypes = [C.c_void_p]
            fn(self.h_)
            self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state writes the caller does directly in the reference (main.cpp:59-79)
    def set_num_particles(self, n):
        fn = self._fn("set_num_particles"); fn.argtypes = [C.c_void_p, C.c_int]; fn(self.h_, n)
        self.num_particles = n

    def set_rest_density(self, d):
        fn = self._fn("set_rest_density"); fn.argtypes = [C.c_void_p, C.c_float]; fn(self.h_, d)
        self.particle_rest_density = _f32(d)

    # ---- phases
    def integrate(self, dt, g):
        fn = self._fn("integrate"); fn.argtypes = [C.c_void_p, C.c_float, C.c_float]; fn(self.h_, dt, g)

    def push_apart(self, iters):
        fn = self._fn("push_apart"); fn.argtypes = [C.c_void_p, C.c_int]; fn(self.h_, iters)

    def bin_particles(self):
        """binning half of pushParticlesApart only (flip_fluid.h:169-198); port back-end only."""
        assert self.kind == "port", "the reference has no separate binning entry point"
        fn = self._fn("bin"); fn.argtypes = [C.c_void_p]; fn(self.h_)

    def separate(self, iters):
        """the serial sweeps of flip_fluid.h:201-250 on the existing bins; port back-end only."""
        assert self.kind == "port"
        fn = self._fn("separate"); fn.argtypes = [C.c_void_p, C.c_int]; fn(self.h_, iters)

    def separate_jacobi(self, iters, with_colors=False):
        """CPU model of the product's parallel separation schedule (NOT the reference's order)."""
        assert self.kind == "port"
        fn = self._fn("separate_jacobi"); fn.argtypes = [C.c_void_p, C.c_int, C.c_int]; fn(self.h_, iters, int(with_colors))

    def collide(self, ox, oy, ovx, ovy, orad):
        fn = self._fn("collide"); fn.argtypes = [C.c_void_p] + [C.c_float] * 5; fn(self.h_, ox, oy, ovx, ovy, orad)

    def transfer(self, to_grid, flip_ratio=0.5):
        fn = self._fn("transfer"); fn.argtypes = [C.c_void_p, C.c_int, C.c_float]; fn(self.h_, int(to_grid), flip_ratio)

    def update_density(self):
        fn = self._fn("density"); fn.argtypes = [C.c_void_p]; fn(self.h_)
        self._refresh_dims()

    def solve(self, iters, dt, omega, drift=True):
        fn = self._fn("solve"); fn.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int]
        fn(self.h_, iters, dt, omega, int(drift))

    def particle_colors(self):
        fn = self._fn("particle_colors"); fn.argtypes = [C.c_void_p]; fn(self.h_)

    def cell_colors(self):
        fn = self._fn("cell_colors"); fn.argtypes = [C.c_void_p]; fn(self.h_)

    def simulate(self, dt, gravity, flip_ratio, num_pressure_iters, num_particle_iters, over_relaxation,
                 compensate_drift, separate_particles, ox, oy, ovx, ovy, orad, n_steps=1, timed=False):
        """FlipFluid::simulate (flip_fluid.h:560-581) n_steps times; timed=True returns ns per phase."""
        fn = self._fn("simulate_timed")
        fn.argtypes = ([C.c_void_p] + [C.c_float] * 3 + [C.c_int] * 2 + [C.c_float] + [C.c_int] * 2 +
                       [C.c_float] * 5 + [C.c_int, C.POINTER(C.c_double)])
        ph = (C.c_double * 8)()
        fn(self.h_, dt, gravity, flip_ratio, num_pressure_iters, num_particle_iters, over_relaxation,
           int(compensate_drift), int(separate_particles), ox, oy, ovx, ovy, orad, n_steps, ph)
        self._refresh_dims()
        return list(ph) if timed else None


class EulerOracle:
    """CPU Eulerian smoke solver (fluid.h:98-259), y-fastest arrays as numpy views."""

    def __init__(self, x, y, density, h, kind=None):
        kind = kind or best_kind("euler")
        self.kind = kind
        self.lib = C.CDLL(_PATHS[("euler", kind)])
        self._p = _PREFIX[("euler", kind)]
        fn = self._fn("create"); fn.restype = C.c_void_p; fn.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float]
        self.h_ = C.c_void_p(fn(x, y, density, h))
        ints = (C.c_int * 3)(); fl = (C.c_float * 3)()
        d = self._fn("dims"); d.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)]
        d(self.h_, ints, fl)
        self.num_x, self.num_y, self.num_cells = list(ints)
        self.density, self.h, self.h1 = [_f32(v) for v in fl]
        ptr = self._fn("ptr"); ptr.restype = C.c_void_p; ptr.argtypes = [C.c_void_p, C.c_char_p]
        for name in ("u", "v", "new_u", "new_v", "pressure", "m", "new_m"):
            a = np.ctypeslib.as_array(C.cast(ptr(self.h_, name.encode()), C.POINTER(C.c_float)), shape=(self.num_cells,))
            setattr(self, name, a.reshape(self.num_x, self.num_y))
        a = np.ctypeslib.as_array(C.cast(ptr(self.h_, b"solid"), C.POINTER(C.c_ubyte)), shape=(self.num_cells,))
        self.solid = a.reshape(self.num_x, self.num_y)

    def _fn(self, name):
        return getattr(self.lib, self._p + name)

    def close(self):
        if self.h_:
            fn = self._fn("destroy"); fn.argtypes = [C.c_void_p]; fn(self.h_); self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self, iters, dt):
        fn = self._fn("solve"); fn.argtypes = [C.c_void_p, C.c_int, C.c_float]; fn(self.h_, iters, dt)

    def extrapolate(self):
        fn = self._fn("extrapolate"); fn.argtypes = [C.c_void_p]; fn(self.h_)

    def advect_vel(self, dt):
        fn = self._fn("advect_vel"); fn.argtypes = [C.c_void_p, C.c_float]; fn(self.h_, dt)

    def advect_smoke(self, dt):
        fn = self._fn("advect_smoke"); fn.argtypes = [C.c_void_p, C.c_float]; fn(self.h_, dt)

    def sample(self, x, y, field):
        fn = self._fn("sample"); fn.restype = C.c_float; fn.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int]
        return fn(self.h_, x, y, field)

    def set_obstacle(self, x, y, r, vx=0.0, vy=0.0):
        """FluidUI::setObstacle (fluid_ui.h:24-62): walls + solid disc carrying velocity (vx, vy)."""
        fn = self._fn("set_obstacle")
        fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]
        fn(self.h_, int(x), int(y), int(r), vx, vy)

    def set_ftz_daz(self, on):
        """SURVEY Q17 timing aid: flush-to-zero + denormals-are-zero on the calling thread; returns the old setting."""
        fn = self._fn("set_ftz_daz"); fn.argtypes = [C.c_int]; fn.restype = C.c_int
        return fn(1 if on else 0)

    def step(self, iters, dt, n_steps=1, timed=False):
        """solve -> extrapolate -> advectVel -> advectSmoke (fluid_ui.h:107-111)."""
        fn = self._fn("step"); fn.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.POINTER(C.c_double)]
        ph = (C.c_double * 2)()
        fn(self.h_, iters, dt, n_steps, ph)
        return list(ph) if timed else None


class Fluid3DOracle:
    """CPU 3D Eulerian smoke solver (eulerian_fluid/src/fluid3d.h:135-367); arrays are numpy views shaped (num_z, num_y, num_x)
    -- x fastest like ix(i,j,k) = i + num_x*j + num_y*num_x*k (:131-133)."""

    def __init__(self, x, y, z, density, h, kind=None):
        kind = kind or best_kind("fluid3d")
        self.kind = kind
        self.lib = C.CDLL(_PATHS[("fluid3d", kind)])
        self._p = _PREFIX[("fluid3d", kind)]
        fn = self._fn("create"); fn.restype = C.c_void_p; fn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]
        self.h_ = C.c_void_p(fn(x, y, z, density, h))
        ints = (C.c_int * 4)(); fl = (C.c_float * 2)()
        d = self._fn("dims"); d.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)]
        d(self.h_, ints, fl)
        self.num_x, self.num_y, self.num_z, self.num_cells = list(ints)
        self.density, self.h = [_f32(v) for v in fl]
        ptr = self._fn("ptr"); ptr.restype = C.c_void_p; ptr.argtypes = [C.c_void_p, C.c_char_p]
        shape = (self.num_z, self.num_y, self.num_x)
        for name in ("u", "v", "w", "new_u", "new_v", "new_w", "pressure", "m", "new_m"):
            a = np.ctypeslib.as_array(C.cast(ptr(self.h_, name.encode()), C.POINTER(C.c_float)), shape=(self.num_cells,))
            setattr(self, name, a.reshape(shape))
        a = np.ctypeslib.as_array(C.cast(ptr(self.h_, b"solid"), C.POINTER(C.c_ubyte)), shape=(self.num_cells,))
        self.solid = a.reshape(shape)

    def _fn(self, name):
        return getattr(self.lib, self._p + name)

    def close(self):
        if self.h_:
            fn = self._fn("destroy"); fn.argtypes = [C.c_void_p]; fn(self.h_); self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self, iters, dt):
        fn = self._fn("solve"); fn.argtypes = [C.c_void_p, C.c_int, C.c_float]; fn(self.h_, iters, dt)

    def extrapolate(self):
        fn = self._fn("extrapolate"); fn.argtypes = [C.c_void_p]; fn(self.h_)

    def advect_vel(self, dt):
        fn = self._fn("advect_vel"); fn.argtypes = [C.c_void_p, C.c_float]; fn(self.h_, dt)

    def advect_smoke(self, dt):
        fn = self._fn("advect_smoke"); fn.argtypes = [C.c_void_p, C.c_float]; fn(self.h_, dt)

    def sample(self, x, y, z, field):
        fn = self._fn("sample"); fn.restype = C.c_float; fn.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int]
        return fn(self.h_, x, y, z, field)

    def step(self, iters, dt, n_steps=1):
        fn = self._fn("step"); fn.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int]; fn(self.h_, iters, dt, n_steps)


def make_fluid3d_scene(x=30, y=22, z=26, oracle_kind=None, seed=7):
    """A wind-tunnel-like 3D scene in the spirit of fluid_ui.h:64-89: inflow u at i=1, solid walls on every face except the
    outflow face, a solid ball, a smoke inlet patch, plus a seeded velocity perturbation so that every term of the advection is
    exercised.  (fluid3d.h has no UI shell; num_z >= num_y keeps the reference's k < num_y loop bound in range, :311.)"""
    o = Fluid3DOracle(x, y, z, 1000.0, 1 / 50.0, kind=oracle_kind)
    nx, ny, nz = o.num_x, o.num_y, o.num_z
    rng = np.random.default_rng(seed)
    o.solid[:] = 0
    o.solid[:, :, 0] = 1; o.solid[:, 0, :] = 1; o.solid[:, -1, :] = 1; o.solid[0, :, :] = 1; o.solid[-1, :, :] = 1
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    ball = (i - nx // 3) ** 2 + (j - ny // 2) ** 2 + (k - nz // 2) ** 2 < (min(nx, ny, nz) // 5) ** 2
    ball[:, :, 0] = ball[:, :, -1] = False
    o.solid[ball] = 1
    o.u[:] = (0.3 * rng.standard_normal(o.u.shape)).astype(np.float32)
    o.v[:] = (0.3 * rng.standard_normal(o.u.shape)).astype(np.float32)
    o.w[:] = (0.3 * rng.standard_normal(o.u.shape)).astype(np.float32)
    o.u[:, :, 1] = 2.0
    for a in (o.u, o.v, o.w):
        a[o.solid != 0] = 0.0
    o.m[:] = 1.0
    o.m[nz // 2 - 2:nz // 2 + 3, ny // 2 - 2:ny // 2 + 3, 0:2] = 0.0
    return o


# --------------------------------------------------------------------------------------
# Scenes: independent numpy-fp32 restatement of the reference UI shells (used to cross-check
# the product's own scene builders in simsandgames_b200/host/scenes.hpp).
# --------------------------------------------------------------------------------------
def flip_tank_layout(tank_width, tank_height, spacing, radius, rel_w=0.6, rel_h=0.8):
    """Particle lattice of flip_fluid/src/main.cpp:47-66 in fp32 (dy goes through double, :49)."""
    f = _f32
    h, r = f(spacing), f(radius)
    dx = f(2) * r
    dy = f(np.sqrt(3.0) / 2 * float(dx))
    num_x = int((f(rel_w) * f(tank_width) - f(2) * h - f(2) * r) / dx)
    num_y = int((f(rel_h) * f(tank_height) - f(2) * h - f(2) * r) / dy)
    i = np.arange(num_x, dtype=np.float32)[:, None]
    j = np.arange(num_y)[None, :]
    x = ((h + r) + dx * i) + np.where(j % 2 == 0, f(0), r).astype(np.float32)
    y = (h + r) + dy * j.astype(np.float32)
    pos = np.empty((num_x, num_y, 2), np.float32)
    pos[..., 0] = x
    pos[..., 1] = np.broadcast_to(y, (num_x, num_y))
    return pos.reshape(-1), num_x, num_y


def flip_obstacle_mask(nx, ny, h, ox, oy, orad):
    """Cells that flip_fluid/src/main.cpp:106-122 turns solid (Q10: loops stop at n-2)."""
    f = _f32
    i = np.arange(nx)[None, :].astype(np.float32)
    j = np.arange(ny)[:, None].astype(np.float32)
    dx = f(h) * (f(.5) + i) - f(ox)
    dy = f(h) * (f(.5) + j) - f(oy)
    inside = dx * dx + dy * dy < f(orad) * f(orad)
    rng = np.zeros((ny, nx), bool)
    rng[1:ny - 2, 1:nx - 2] = True
    return inside & rng, rng


def make_flip_scene(kind="default", N=None, oracle_kind=None, with_obstacle=None):
    """Build a FlipOracle initialised like the reference UI shell.

    kind="default": flip_fluid/src/main.cpp:33-79 (640x480 window, res=100) -> S0.
    kind="synthetic": SURVEY 8d rule for S2/S3/S4 (sp=3/N, W=H=(N-.5)sp, r=.19sp), obstacle optional.
    Returns (oracle, params dict with the simulate() arguments).
    """
    f = _f32
    if kind == "default":
        sim_h = f(3)
        c_scale = f(480) / sim_h
        sim_w = f(640) / c_scale
        W, H = sim_w, sim_h
        sp = H / f(100)
        r = f(.3) * sp
        dt = f(1) / f(60)
        obstacle = (f(3), f(2), f(.15))
        iters = 50
    else:
        sp = f(3) / f(N)
        W = H = (f(N) - f(.5)) * sp
        r = f(.19) * sp
        dt = (f(1) / f(60)) * (f(100) / f(N))
        obstacle = (f(.75) * W, f(.67) * H, f(.05) * W) if with_obstacle else None
        iters = 200 if with_obstacle else 50
    pos, num_x, num_y = flip_tank_layout(W, H, sp, r)
    P = num_x * num_y
    o = FlipOracle(1000.0, W, H, sp, r, P, kind=oracle_kind)
    o.set_num_particles(P)
    o.particle_pos[:] = pos
    s = np.ones((o.f_num_y, o.f_num_x), np.float32)
    s[0, :] = 0; s[-1, :] = 0; s[:, 0] = 0; s[:, -1] = 0
    ox = oy = f(0); orad = f(0)
    if obstacle is not None:
        ox, oy, orad = obstacle
        inside, rng = flip_obstacle_mask(o.f_num_x, o.f_num_y, o.h, ox, oy, orad)
        s[rng] = 1
        s[inside] = 0
    o.s[:] = s.reshape(-1)
    params = dict(dt=float(dt), gravity=9.81, flip_ratio=0.9, num_pressure_iters=iters, num_particle_iters=2,
                  over_relaxation=1.9, compensate_drift=True, separate_particles=True,
                  ox=float(ox), oy=float(oy), ovx=0.0, ovy=0.0, orad=float(orad))
    geom = dict(density=1000.0, width=float(W), height=float(H), spacing=float(sp), radius=float(r),
                max_particles=P, num_x=num_x, num_y=num_y)
    return o, params, geom


def euler_set_obstacle(o, x, y, r, vx=0.0, vy=0.0):
    """numpy restatement of eulerian_fluid/src/fluid_ui.h:25-62 (integer disc, left/top/bottom walls); the C
    restatements are EulerOracle.set_obstacle, cross-checked against this one in tests/test_oracle_pin.py."""
    o.solid[:] = 0
    o.solid[0, :] = 1
    o.solid[:, 0] = 1
    o.solid[:, -1] = 1
    i = np.arange(o.num_x)[:, None]
    j = np.arange(o.num_y)[None, :]
    disc = ((i - x) ** 2 + (j - y) ** 2 < r * r)
    disc[0, :] = False; disc[-1, :] = False; disc[:, 0] = False; disc[:, -1] = False
    o.solid[disc] = 1
    o.m[disc] = 1.0
    o.u[disc] = vx
    o.v[disc] = vy


def make_euler_scene(x=96, y=72, oracle_kind=None):
    """Wind tunnel of eulerian_fluid/src/fluid_ui.h:64-89, scaled to an x-by-y interior."""
    o = EulerOracle(x, y, 1000.0, 1 / 100.0, kind=oracle_kind)
    o.u[1, :] = 2.0
    euler_set_obstacle(o, o.num_x // 3, o.num_y // 2, o.num_x // 8)
    pipe_h = o.num_y // 6
    lo = o.num_y // 2 - pipe_h // 2
    hi = o.num_y // 2 + pipe_h // 2
    o.m[0, lo:hi + 1] = 0.0
    return o
