"""ctypes binding of libflipgpu.so (include/flipgpu.h) with the reference's member names.

FlipFluidGPU mirrors struct FlipFluid (reference flip_fluid/src/flip_fluid.h:13-582):
same constructor arguments (:62-65), same simulate() signature (:560-565), same public
scalars (f_num_x, h, particle_radius, num_particles, ...), and fIX/pIX.  The public raw arrays
of the reference (u, v, s, particle_pos, ...) live in HBM here, so they are exchanged with
`upload(name, array)` / `readback(name)` in the reference's layout and index order.
FluidGPU mirrors class Fluid (reference eulerian_fluid/src/fluid.h:19-260).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

LEX_WAVEFRONT = 0   # FlipSolverOrder of include/flipgpu.h: 0 = the reference order
RED_BLACK = 1
LEX_DIAGONAL = 2
LEX_SYSTOLIC = 3

FLIP_FIELDS = {  # name -> (enum value, numpy dtype)   order == FlipField in flipgpu.h
    "u": (0, np.float32), "v": (1, np.float32), "du": (2, np.float32), "dv": (3, np.float32),
    "prev_u": (4, np.float32), "prev_v": (5, np.float32), "p": (6, np.float32), "s": (7, np.float32),
    "cell_type": (8, np.int32), "cell_color": (9, np.float32),
    "particle_pos": (10, np.float32), "particle_vel": (11, np.float32), "particle_color": (12, np.float32),
    "particle_density": (13, np.float32),
    "num_cell_particles": (14, np.int32), "first_cell_particle": (15, np.int32), "cell_particle_ids": (16, np.int32),
}
EULER_FIELDS = {"u": (0, np.float32), "v": (1, np.float32), "new_u": (2, np.float32), "new_v": (3, np.float32),
                "pressure": (4, np.float32), "m": (5, np.float32), "new_m": (6, np.float32), "solid": (7, np.uint8)}
FLIP_PHASES = ["integrate_bin", "separate", "collide_mark", "p2g", "solve", "g2p", "colors"]


class FlipGpuError(RuntimeError):
    pass


class _FlipParams(C.Structure):
    _fields_ = [("density", C.c_float), ("width", C.c_float), ("height", C.c_float), ("spacing", C.c_float),
                ("particle_radius", C.c_float), ("max_particles", C.c_int), ("device", C.c_int)]


class _FlipDims(C.Structure):
    _fields_ = [("f_num_x", C.c_int), ("f_num_y", C.c_int), ("f_num_cells", C.c_int),
                ("p_num_x", C.c_int), ("p_num_y", C.c_int), ("p_num_cells", C.c_int),
                ("max_particles", C.c_int), ("num_particles", C.c_int),
                ("density", C.c_float), ("h", C.c_float), ("f_inv_spacing", C.c_float),
                ("particle_radius", C.c_float), ("p_inv_spacing", C.c_float)]


class _FlipStepParams(C.Structure):
    _fields_ = [("dt", C.c_float), ("gravity", C.c_float), ("flip_ratio", C.c_float),
                ("num_pressure_iters", C.c_int), ("num_particle_iters", C.c_int),
                ("over_relaxation", C.c_float), ("compensate_drift", C.c_int), ("separate_particles", C.c_int),
                ("obstacle_x", C.c_float), ("obstacle_y", C.c_float), ("obstacle_vel_x", C.c_float),
                ("obstacle_vel_y", C.c_float), ("obstacle_radius", C.c_float),
                ("solver_order", C.c_int), ("update_colors", C.c_int)]


class _EulerDims(C.Structure):
    _fields_ = [("num_x", C.c_int), ("num_y", C.c_int), ("num_cells", C.c_int),
                ("density", C.c_float), ("h", C.c_float), ("h1", C.c_float)]


def lib_path():
    # FLIPGPU_LIB: an alternative build of the same library (A/B of compile-time variants, scripts/ only)
    return os.environ.get("FLIPGPU_LIB") or os.path.join(_HERE, "libflipgpu.so")


def build_library(verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> simsandgames_b200/libflipgpu.so (in-tree)."""
    r = subprocess.run(["make", "-j8", "-C", os.path.join(_HERE, "csrc")], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise FlipGpuError("building libflipgpu.so failed:\n" + (r.stdout or "") + (r.stderr or ""))


def load_library():
    """Load the C-ABI library.  Raises if it has not been built: there is no fallback path."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise FlipGpuError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(simsandgames_b200 has no CPU fallback)")
    lib = C.CDLL(path)
    lib.flipgpu_last_error.restype = C.c_char_p
    _LIB = lib
    return lib


def _check(rc, what):
    if rc != 0:
        msg = load_library().flipgpu_last_error().decode(errors="replace")
        raise FlipGpuError(f"{what} failed (code {rc}): {msg}")


def _as_c(arr, dtype):
    a = np.ascontiguousarray(arr, dtype=dtype)
    return a, a.ctypes.data_as(C.c_void_p)


class FlipFluidGPU:
    """FlipFluid(d, width, height, sp, r, m) -- flip_fluid.h:62-65 -- with its state in B200 HBM."""

    def __init__(self, density, width, height, spacing, particle_radius, max_particles, device=-1):
        self._lib = load_library()
        self._h = C.c_void_p()
        pr = _FlipParams(density, width, height, spacing, particle_radius, max_particles, device)
        _check(self._lib.flip_create(C.byref(pr), C.byref(self._h)), "flip_create")
        self._refresh()

    def _refresh(self):
        d = _FlipDims()
        _check(self._lib.flip_get_dims(self._h, C.byref(d)), "flip_get_dims")
        for name, _ in _FlipDims._fields_:
            v = getattr(d, name)
            setattr(self, "_" + name if name == "num_particles" else name, np.float32(v) if isinstance(v, float) else v)

    # ---- reference members
    @property
    def num_particles(self):
        return self._num_particles

    @num_particles.setter
    def num_particles(self, n):  # flip_fluid/src/main.cpp:59 writes this member directly
        _check(self._lib.flip_set_num_particles(self._h, int(n)), "flip_set_num_particles")
        self._num_particles = int(n)

    @property
    def particle_rest_density(self):
        out = C.c_float()
        _check(self._lib.flip_get_scalar(self._h, 0, C.byref(out)), "flip_get_scalar")
        return np.float32(out.value)

    @particle_rest_density.setter
    def particle_rest_density(self, v):
        _check(self._lib.flip_set_rest_density(self._h, C.c_float(v)), "flip_set_rest_density")

    def fIX(self, i, j):  # flip_fluid.h:147-149
        return i + self.f_num_x * j

    def pIX(self, i, j):  # flip_fluid.h:151-153
        return i + self.p_num_x * j

    def _count(self, name):
        Cn, P, H = self.f_num_cells, self._num_particles, self.p_num_cells
        return {"cell_color": 3 * Cn, "particle_pos": 2 * P, "particle_vel": 2 * P, "particle_color": 3 * P,
                "num_cell_particles": H, "first_cell_particle": H + 1, "cell_particle_ids": P}.get(name, Cn)

    def upload(self, name, array):
        fid, dt = FLIP_FIELDS[name]
        a, p = _as_c(np.asarray(array).reshape(-1), dt)
        _check(self._lib.flip_upload(self._h, fid, p, C.c_size_t(a.size)), f"flip_upload({name})")

    def readback(self, name, out=None):
        fid, dt = FLIP_FIELDS[name]
        n = self._count(name)
        if out is None:
            out = np.empty(n, dt)
        _check(self._lib.flip_readback(self._h, fid, out.ctypes.data_as(C.c_void_p), C.c_size_t(n)), f"flip_readback({name})")
        return out

    def readback_into(self, name, ptr, count):
        """readback into a raw host pointer (e.g. pinned torch memory)."""
        fid, _ = FLIP_FIELDS[name]
        _check(self._lib.flip_readback(self._h, fid, C.c_void_p(ptr), C.c_size_t(count)), f"flip_readback({name})")

    def upload_from(self, name, ptr, count):
        fid, _ = FLIP_FIELDS[name]
        _check(self._lib.flip_upload(self._h, fid, C.c_void_p(ptr), C.c_size_t(count)), f"flip_upload({name})")

    def upload_from_async(self, name, ptr, count):
        """flip_upload_async: the copy is only queued; the (pinned) host buffer must stay unchanged until the next sync"""
        fid, _ = FLIP_FIELDS[name]
        _check(self._lib.flip_upload_async(self._h, fid, C.c_void_p(ptr), C.c_size_t(count)), f"flip_upload_async({name})")

    def readback_pos_async(self, ptr, count):
        """particle_pos -> (pinned) host pointer on the copy stream; returns at once, join with readback_wait()."""
        _check(self._lib.flip_readback_pos_async(self._h, C.c_void_p(ptr), C.c_size_t(count)), "flip_readback_pos_async")

    def readback_wait(self):
        _check(self._lib.flip_readback_wait(self._h), "flip_readback_wait")

    def set_obstacle(self, x, y, vx=0.0, vy=0.0, radius=0.15):
        """device equivalent of FluidUI::setObstacle, flip_fluid/src/main.cpp:95-126"""
        _check(self._lib.flip_set_obstacle(self._h, *[C.c_float(t) for t in (x, y, vx, vy, radius)]), "flip_set_obstacle")

    def simulate(self, dt, gravity, flip_ratio, num_pressure_iters, num_particle_iters, over_relaxation,
                 compensate_drift, separate_particles, obstacle_x, obstacle_y, obstacle_vel_x, obstacle_vel_y,
                 obstacle_radius, n_steps=1, solver_order=LEX_WAVEFRONT, update_colors=True):
        """FlipFluid::simulate, flip_fluid.h:560-565 (the first 13 arguments are the reference's)."""
        sp = _FlipStepParams(dt, gravity, flip_ratio, num_pressure_iters, num_particle_iters, over_relaxation,
                             int(compensate_drift), int(separate_particles), obstacle_x, obstacle_y,
                             obstacle_vel_x, obstacle_vel_y, obstacle_radius, solver_order, int(update_colors))
        _check(self._lib.flip_step(self._h, C.byref(sp), int(n_steps)), "flip_step")

    # ---- single phases (same-state replay)
    def integrateParticles(self, dt, gravity):
        _check(self._lib.flip_phase_integrate(self._h, C.c_float(dt), C.c_float(gravity)), "flip_phase_integrate")

    def binParticles(self):
        _check(self._lib.flip_phase_bin(self._h), "flip_phase_bin")

    def separateParticles(self, num_iter, update_colors=False):
        _check(self._lib.flip_phase_separate(self._h, int(num_iter), int(update_colors)), "flip_phase_separate")

    def pushParticlesApart(self, num_iter, update_colors=False):  # flip_fluid.h:165-251 = bin + sweeps
        self.binParticles()
        self.separateParticles(num_iter, update_colors)

    def handleParticleCollisions(self, ox, oy, ovx, ovy, orad):
        _check(self._lib.flip_phase_collide(self._h, *[C.c_float(t) for t in (ox, oy, ovx, ovy, orad)]), "flip_phase_collide")

    def transferVelocitiesToGrid(self):  # transferVelocities(true) + updateParticleDensity
        _check(self._lib.flip_phase_p2g(self._h), "flip_phase_p2g")

    def solveIncompressibility(self, num_iters, dt, over_relaxation, compensate_drift=True, solver_order=LEX_WAVEFRONT):
        _check(self._lib.flip_phase_solve(self._h, int(num_iters), C.c_float(dt), C.c_float(over_relaxation),
                                          int(compensate_drift), int(solver_order)), "flip_phase_solve")

    def transferVelocitiesToParticles(self, flip_ratio):
        _check(self._lib.flip_phase_g2p(self._h, C.c_float(flip_ratio)), "flip_phase_g2p")

    def updateColors(self):
        _check(self._lib.flip_phase_colors(self._h), "flip_phase_colors")

    def set_exact_order(self, on=True):
        """reference order of every float reduction (serial-order separation, caller-order P2G, serial latch): parity mode"""
        _check(self._lib.flip_set_exact_order(self._h, int(on)), "flip_set_exact_order")

    def exact_order_info(self):
        a, b = C.c_int(), C.c_int()
        _check(self._lib.flip_get_exact_order_info(self._h, C.byref(a), C.byref(b)), "flip_get_exact_order_info")
        return a.value, b.value

    # ---- diagnostics
    def residual(self):
        mx, rms = C.c_float(), C.c_float()
        _check(self._lib.flip_get_scalar(self._h, 1, C.byref(mx)), "flip_get_scalar")
        _check(self._lib.flip_get_scalar(self._h, 2, C.byref(rms)), "flip_get_scalar")
        return float(mx.value), float(rms.value)

    def synchronize(self):
        _check(self._lib.flip_synchronize(self._h), "flip_synchronize")

    @property
    def stream(self):
        s = C.c_void_p()
        _check(self._lib.flip_get_stream(self._h, C.byref(s)), "flip_get_stream")
        return s.value or 0

    def enable_timings(self, on=True):
        _check(self._lib.flip_enable_timings(self._h, int(on)), "flip_enable_timings")

    def reset_timings(self):
        _check(self._lib.flip_reset_timings(self._h), "flip_reset_timings")

    def timings(self):
        ms = (C.c_float * len(FLIP_PHASES))()
        _check(self._lib.flip_get_timings(self._h, ms, len(FLIP_PHASES)), "flip_get_timings")
        return dict(zip(FLIP_PHASES, list(ms)))

    @property
    def launch_count(self):
        n = C.c_longlong()
        _check(self._lib.flip_get_launch_count(self._h, C.byref(n)), "flip_get_launch_count")
        return n.value

    def set_graph_mode(self, on=True):
        _check(self._lib.flip_set_graph_mode(self._h, int(on)), "flip_set_graph_mode")

    @property
    def graph_replays(self):
        n = C.c_longlong()
        _check(self._lib.flip_get_graph_replays(self._h, C.byref(n)), "flip_get_graph_replays")
        return n.value

    def solver_profile(self):
        """FLIPGPU_SOR_PROFILE=1: SM cycles the lexicographic solver spent per phase since the last call (8 counters)"""
        out = (C.c_ulonglong * 8)()
        _check(self._lib.flip_debug_solver_profile(self._h, out), "flip_debug_solver_profile")
        return list(out)

    def close(self):
        if self._h:
            self._lib.flip_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FluidGPU:
    """Fluid(x, y, d, h) -- eulerian_fluid/src/fluid.h:41 -- with its state in B200 HBM (y-fastest arrays)."""
    U_FIELD, V_FIELD, S_FIELD = 0, 1, 2  # fluid.h:33-37

    def __init__(self, x, y, density, h, device=-1):
        self._lib = load_library()
        self._h = C.c_void_p()
        _check(self._lib.euler_create(int(x), int(y), C.c_float(density), C.c_float(h), int(device), C.byref(self._h)),
               "euler_create")
        d = _EulerDims()
        _check(self._lib.euler_get_dims(self._h, C.byref(d)), "euler_get_dims")
        self._num_x, self._num_y, self._num_cells = d.num_x, d.num_y, d.num_cells
        self.density, self.h, self.h1 = np.float32(d.density), np.float32(d.h), np.float32(d.h1)
        self.obstacle_x = self.obstacle_y = 0  # FluidUI's members, fluid_ui.h:17-18

    def getNumX(self):
        return self._num_x

    def getNumY(self):
        return self._num_y

    def getNumCells(self):
        return self._num_cells

    def ix(self, i, j):  # fluid.h:94-96
        return i * self._num_y + j

    def upload(self, name, array):
        fid, dt = EULER_FIELDS[name]
        a, p = _as_c(np.asarray(array).reshape(-1), dt)
        _check(self._lib.euler_upload(self._h, fid, p, C.c_size_t(a.size)), f"euler_upload({name})")

    def readback(self, name):
        fid, dt = EULER_FIELDS[name]
        out = np.empty(self._num_cells, dt)
        _check(self._lib.euler_readback(self._h, fid, out.ctypes.data_as(C.c_void_p), C.c_size_t(out.size)), f"euler_readback({name})")
        return out.reshape(self._num_x, self._num_y)

    def readback_into(self, name, ptr, count):
        fid, _ = EULER_FIELDS[name]
        _check(self._lib.euler_readback(self._h, fid, C.c_void_p(ptr), C.c_size_t(count)), f"euler_readback({name})")

    def upload_from(self, name, ptr, count):
        fid, _ = EULER_FIELDS[name]
        _check(self._lib.euler_upload(self._h, fid, C.c_void_p(ptr), C.c_size_t(count)), f"euler_upload({name})")

    def solveIncompressibility(self, num_iter, dt, solver_order=LEX_WAVEFRONT):
        _check(self._lib.euler_project(self._h, int(num_iter), C.c_float(dt), int(solver_order)), "euler_project")

    def extrapolate(self):
        _check(self._lib.euler_extrapolate(self._h), "euler_extrapolate")

    def advectVel(self, dt):
        _check(self._lib.euler_advect_vel(self._h, C.c_float(dt)), "euler_advect_vel")

    def advectSmoke(self, dt):
        _check(self._lib.euler_advect_smoke(self._h, C.c_float(dt)), "euler_advect_smoke")

    def step(self, num_iter, dt, n_steps=1, solver_order=LEX_WAVEFRONT):
        """the caller sequence of eulerian_fluid/src/fluid_ui.h:107-111"""
        _check(self._lib.euler_step(self._h, int(num_iter), C.c_float(dt), int(solver_order), int(n_steps)), "euler_step")

    def setObstacle(self, x, y, r, reset=True, dt=1.0):
        """FluidUI::setObstacle(x, y, r, reset, dt) -- eulerian_fluid/src/fluid_ui.h:24-62 -- on the device."""
        vx = vy = np.float32(0)
        if not reset:  # :39-43, velocity = displacement of the obstacle since the previous call
            vx = np.float32(int(x) - self.obstacle_x) / np.float32(dt)
            vy = np.float32(int(y) - self.obstacle_y) / np.float32(dt)
        self.obstacle_x, self.obstacle_y = int(x), int(y)
        _check(self._lib.euler_set_obstacle(self._h, int(x), int(y), int(r), C.c_float(vx), C.c_float(vy)), "euler_set_obstacle")

    def sampleField(self, xs, ys, field):
        xs = np.ascontiguousarray(xs, np.float32).reshape(-1)
        ys = np.ascontiguousarray(ys, np.float32).reshape(-1)
        out = np.empty_like(xs)
        _check(self._lib.euler_sample_field(self._h, int(field), xs.ctypes.data_as(C.POINTER(C.c_float)),
                                            ys.ctypes.data_as(C.POINTER(C.c_float)),
                                            out.ctypes.data_as(C.POINTER(C.c_float)), C.c_size_t(xs.size)), "euler_sample_field")
        return out

    def residual(self):
        mx, rms = C.c_float(), C.c_float()
        _check(self._lib.euler_get_residual(self._h, C.byref(mx), C.byref(rms)), "euler_get_residual")
        return float(mx.value), float(rms.value)

    def synchronize(self):
        _check(self._lib.euler_synchronize(self._h), "euler_synchronize")

    @property
    def stream(self):
        s = C.c_void_p()
        _check(self._lib.euler_get_stream(self._h, C.byref(s)), "euler_get_stream")
        return s.value or 0

    @property
    def launch_count(self):
        n = C.c_longlong()
        _check(self._lib.euler_get_launch_count(self._h, C.byref(n)), "euler_get_launch_count")
        return n.value

    def close(self):
        if self._h:
            self._lib.euler_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


FLUID3D_FIELDS = {"u": (0, np.float32), "v": (1, np.float32), "w": (2, np.float32), "new_u": (3, np.float32), "new_v": (4, np.float32),
                  "new_w": (5, np.float32), "pressure": (6, np.float32), "m": (7, np.float32), "new_m": (8, np.float32), "solid": (9, np.uint8)}


class _Fluid3DDims(C.Structure):
    _fields_ = [("num_x", C.c_int), ("num_y", C.c_int), ("num_z", C.c_int), ("num_cells", C.c_int), ("density", C.c_float), ("h", C.c_float)]


class Fluid3DGPU:
    """Fluid3D(x, y, z, d, h) -- eulerian_fluid/src/fluid3d.h:40 -- with its state in B200 HBM; arrays cross the boundary in the
    reference layout ix(i,j,k) = i + num_x*j + num_y*num_x*k, i.e. numpy shape (num_z, num_y, num_x)."""
    U_FIELD, V_FIELD, W_FIELD, S_FIELD = 0, 1, 2, 3  # fluid3d.h:31-36

    def __init__(self, x, y, z, density, h, device=-1):
        self._lib = load_library()
        self._h = C.c_void_p()
        _check(self._lib.fluid3d_create(int(x), int(y), int(z), C.c_float(density), C.c_float(h), int(device), C.byref(self._h)), "fluid3d_create")
        d = _Fluid3DDims()
        _check(self._lib.fluid3d_get_dims(self._h, C.byref(d)), "fluid3d_get_dims")
        self.num_x, self.num_y, self.num_z, self.num_cells = d.num_x, d.num_y, d.num_z, d.num_cells
        self.density, self.h = np.float32(d.density), np.float32(d.h)

    def ix(self, i, j, k):  # fluid3d.h:131-133
        return i + self.num_x * j + self.num_y * self.num_x * k

    def upload(self, name, array):
        fid, dt = FLUID3D_FIELDS[name]
        a, p = _as_c(np.asarray(array).reshape(-1), dt)
        _check(self._lib.fluid3d_upload(self._h, fid, p, C.c_size_t(a.size)), f"fluid3d_upload({name})")

    def readback(self, name):
        fid, dt = FLUID3D_FIELDS[name]
        out = np.empty(self.num_cells, dt)
        _check(self._lib.fluid3d_readback(self._h, fid, out.ctypes.data_as(C.c_void_p), C.c_size_t(out.size)), f"fluid3d_readback({name})")
        return out.reshape(self.num_z, self.num_y, self.num_x)

    def solveIncompressibility(self, num_iter, dt, solver_order=LEX_WAVEFRONT):
        _check(self._lib.fluid3d_project(self._h, int(num_iter), C.c_float(dt), int(solver_order)), "fluid3d_project")

    def extrapolate(self):
        _check(self._lib.fluid3d_extrapolate(self._h), "fluid3d_extrapolate")

    def advectVel(self, dt):
        _check(self._lib.fluid3d_advect_vel(self._h, C.c_float(dt)), "fluid3d_advect_vel")

    def advectSmoke(self, dt):
        _check(self._lib.fluid3d_advect_smoke(self._h, C.c_float(dt)), "fluid3d_advect_smoke")

    def step(self, num_iter, dt, n_steps=1, solver_order=LEX_WAVEFRONT):
        _check(self._lib.fluid3d_step(self._h, int(num_iter), C.c_float(dt), int(solver_order), int(n_steps)), "fluid3d_step")

    def sampleField(self, xs, ys, zs, field):
        xs, ys, zs = (np.ascontiguousarray(a, np.float32).reshape(-1) for a in (xs, ys, zs))
        out = np.empty_like(xs)
        P = C.POINTER(C.c_float)
        _check(self._lib.fluid3d_sample_field(self._h, int(field), xs.ctypes.data_as(P), ys.ctypes.data_as(P), zs.ctypes.data_as(P),
                                              out.ctypes.data_as(P), C.c_size_t(xs.size)), "fluid3d_sample_field")
        return out

    def synchronize(self):
        _check(self._lib.fluid3d_synchronize(self._h), "fluid3d_synchronize")

    @property
    def launch_count(self):
        n = C.c_longlong()
        _check(self._lib.fluid3d_get_launch_count(self._h, C.byref(n)), "fluid3d_get_launch_count")
        return n.value

    def close(self):
        if self._h:
            self._lib.fluid3d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


SHARD_FIELDS = {"vel_fast": (0, np.float32), "vel_slow": (1, np.float32), "p": (2, np.float32), "s": (3, np.float32),
                "density": (4, np.float32), "cell_type": (5, np.int32)}


def nccl_unique_id():
    """128-byte NCCL communicator id (rank 0 creates it, the caller broadcasts it)."""
    buf = (C.c_ubyte * 128)()
    _check(load_library().flipgpu_nccl_unique_id(buf), "flipgpu_nccl_unique_id")
    return bytes(buf)


def slab_rows(ns, rank, nranks):
    """Rows [j0, j1) of rank `rank` when ns rows are cut into nranks slabs (same rule as the library)."""
    j0, j1 = C.c_int(), C.c_int()
    load_library().sorshard_slab(int(ns), int(rank), int(nranks), C.byref(j0), C.byref(j1))
    return j0.value, j1.value


class ShardedProjection:
    """One rank's slab of the red-black SOR projection; arrays are (rows, nf) numpy views of global rows."""

    def __init__(self, nf, ns, x_fast, rank, nranks, unique_id=None, device=-1):
        self._lib = load_library()
        self._h = C.c_void_p()
        idbuf = (C.c_ubyte * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        _check(self._lib.sorshard_create(int(nf), int(ns), int(bool(x_fast)), int(rank), int(nranks), idbuf, int(device), C.byref(self._h)),
               "sorshard_create")
        self.nf, self.ns, self.rank, self.nranks = nf, ns, rank, nranks
        j0, j1, halo = C.c_int(), C.c_int(), C.c_int()
        _check(self._lib.sorshard_get_rows(self._h, C.byref(j0), C.byref(j1), C.byref(halo)), "sorshard_get_rows")
        self.j0, self.j1, self.halo = j0.value, j1.value, halo.value
        self.w0, self.w1 = max(self.j0 - self.halo, 0), min(self.j1 + self.halo, ns)   # window incl. halo rows

    def upload_window(self, name, global_array):
        """Copy this rank's window rows [j0-halo, j1+halo) out of a full (ns, nf) host array."""
        fid, dt = SHARD_FIELDS[name]
        a = np.ascontiguousarray(np.asarray(global_array).reshape(self.ns, self.nf)[self.w0:self.w1], dtype=dt)
        _check(self._lib.sorshard_upload_rows(self._h, fid, a.ctypes.data_as(C.c_void_p), self.w0, self.w1 - self.w0), f"sorshard_upload_rows({name})")

    def upload_rows(self, name, rows, j_start):
        fid, dt = SHARD_FIELDS[name]
        a = np.ascontiguousarray(rows, dtype=dt).reshape(-1, self.nf)
        _check(self._lib.sorshard_upload_rows(self._h, fid, a.ctypes.data_as(C.c_void_p), int(j_start), a.shape[0]), f"sorshard_upload_rows({name})")

    def readback_own(self, name):
        fid, dt = SHARD_FIELDS[name]
        out = np.empty((self.j1 - self.j0, self.nf), dt)
        _check(self._lib.sorshard_readback_rows(self._h, fid, out.ctypes.data_as(C.c_void_p), self.j0, self.j1 - self.j0), f"sorshard_readback_rows({name})")
        return out

    def solve(self, iters, cp, omega, compensate_drift=False, rest_density=0.0):
        _check(self._lib.sorshard_solve(self._h, int(iters), C.c_float(cp), C.c_float(omega), int(compensate_drift), C.c_float(rest_density)),
               "sorshard_solve")

    def synchronize(self):
        _check(self._lib.sorshard_synchronize(self._h), "sorshard_synchronize")

    @property
    def stream(self):
        s = C.c_void_p()
        _check(self._lib.sorshard_get_stream(self._h, C.byref(s)), "sorshard_get_stream")
        return s.value or 0

    def counters(self):
        a, b = C.c_longlong(), C.c_longlong()
        _check(self._lib.sorshard_get_counters(self._h, C.byref(a), C.byref(b)), "sorshard_get_counters")
        return a.value, b.value

    def close(self):
        if self._h:
            self._lib.sorshard_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FlipFluidSharded(FlipFluidGPU):
    """One rank's slab of a slab-sharded FlipFluid.  Grid arrays keep the global shape (only the rank's
    window rows are current); particles are exchanged as (pos, vel, global id) triples."""

    def __init__(self, density, width, height, spacing, particle_radius, max_local_particles, global_particles, rank, nranks,
                 unique_id=None, device=-1, row_starts=None):
        self._lib = load_library()
        self._h = C.c_void_p()
        pr = _FlipParams(density, width, height, spacing, particle_radius, max_local_particles, device)
        idbuf = (C.c_ubyte * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        rows = None
        if row_starts is not None:   # caller-chosen slabs (load balance), else equal slabs
            assert len(row_starts) == nranks + 1
            rows = (C.c_int * (nranks + 1))(*[int(r) for r in row_starts])
        _check(self._lib.flip_create_sharded_rows(C.byref(pr), int(global_particles), int(rank), int(nranks), rows, idbuf, C.byref(self._h)),
               "flip_create_sharded_rows")
        self._refresh()
        self.rank, self.nranks = rank, nranks
        self.j0, self.j1, self.halo, self.band, _ = self.shard_info()

    def shard_info(self):
        a, b, h, g, e = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_longlong()
        _check(self._lib.flip_get_shard_info(self._h, C.byref(a), C.byref(b), C.byref(h), C.byref(g), C.byref(e)), "flip_get_shard_info")
        return a.value, b.value, h.value, g.value, e.value

    def owns(self, pos):
        """mask of the particles (global (P,2) positions) whose cell row lies in this rank's slab"""
        y = np.asarray(pos, np.float32).reshape(-1, 2)[:, 1]
        row = np.clip((np.float32(self.f_inv_spacing) * y).astype(np.int32), 0, self.f_num_y - 1)
        return (row >= self.j0) & (row < self.j1)

    def upload_particles(self, pos, vel, ids):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1)
        ids = np.ascontiguousarray(ids, np.int32).reshape(-1)
        vp = None
        if vel is not None:
            vel = np.ascontiguousarray(vel, np.float32).reshape(-1)
            vp = vel.ctypes.data_as(C.POINTER(C.c_float))
        _check(self._lib.flip_upload_particles(self._h, pos.ctypes.data_as(C.POINTER(C.c_float)), vp,
                                               ids.ctypes.data_as(C.POINTER(C.c_int)), int(ids.size)), "flip_upload_particles")
        self._num_particles = int(ids.size)

    def readback_particles(self):
        cap = self.max_particles
        pos = np.empty(2 * cap, np.float32); vel = np.empty(2 * cap, np.float32); ids = np.empty(cap, np.int32)
        n = C.c_int()
        _check(self._lib.flip_readback_particles(self._h, pos.ctypes.data_as(C.POINTER(C.c_float)), vel.ctypes.data_as(C.POINTER(C.c_float)),
                                                 ids.ctypes.data_as(C.POINTER(C.c_int)), cap, C.byref(n)), "flip_readback_particles")
        k = n.value
        return pos[:2 * k].copy(), vel[:2 * k].copy(), ids[:k].copy()

    def readback_owned_pos_into(self, ptr, capacity):
        """owned particle positions -> raw (pinned) host pointer; returns how many particles"""
        n = C.c_int()
        _check(self._lib.flip_readback_particles(self._h, C.cast(C.c_void_p(ptr), C.POINTER(C.c_float)), None, None, int(capacity), C.byref(n)),
               "flip_readback_particles")
        return n.value

    def upload_particles_from(self, pos_ptr, ids_ptr, n):
        _check(self._lib.flip_upload_particles(self._h, C.cast(C.c_void_p(pos_ptr), C.POINTER(C.c_float)), None,
                                               C.cast(C.c_void_p(ids_ptr), C.POINTER(C.c_int)), int(n)), "flip_upload_particles")
        self._num_particles = int(n)

    def upload_rows_from(self, name, ptr, j_start, n_rows):
        fid, _ = FLIP_FIELDS[name]
        _check(self._lib.flip_upload_rows(self._h, fid, C.c_void_p(ptr), int(j_start), int(n_rows)), f"flip_upload_rows({name})")

    def upload_rows(self, name, rows, j_start):
        fid, dt = FLIP_FIELDS[name]
        a = np.ascontiguousarray(rows, dtype=dt).reshape(-1, self.f_num_x)
        _check(self._lib.flip_upload_rows(self._h, fid, a.ctypes.data_as(C.c_void_p), int(j_start), a.shape[0]), f"flip_upload_rows({name})")

    def readback_own_rows(self, name):
        fid, dt = FLIP_FIELDS[name]
        out = np.empty((self.j1 - self.j0, self.f_num_x), dt)
        _check(self._lib.flip_readback_rows(self._h, fid, out.ctypes.data_as(C.c_void_p), self.j0, self.j1 - self.j0), f"flip_readback_rows({name})")
        return out

    def simulate(self, *args, **kw):
        kw.setdefault("solver_order", RED_BLACK)
        kw.setdefault("update_colors", False)
        super().simulate(*args, **kw)
