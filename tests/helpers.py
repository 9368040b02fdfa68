"""Shared helpers for the parity tests: move state between the CPU oracle and the GPU sim."""
import numpy as np
from oracle import cpu_oracle as co
import simsandgames_b200 as sg

GRID_FIELDS = ["u", "v", "prev_u", "prev_v", "p", "s", "particle_density"]


def oracle_args(params):
    """simulate() kwargs of the product -> kwargs of oracle.simulate"""
    m = {"obstacle_x": "ox", "obstacle_y": "oy", "obstacle_vel_x": "ovx", "obstacle_vel_y": "ovy", "obstacle_radius": "orad"}
    return {m.get(k, k): v for k, v in params.items()}


def product_args(params):
    m = {"ox": "obstacle_x", "oy": "obstacle_y", "ovx": "obstacle_vel_x", "ovy": "obstacle_vel_y", "orad": "obstacle_radius"}
    return {m.get(k, k): v for k, v in params.items()}


def gpu_like(o, geom):
    """A FlipFluidGPU with the same constructor arguments as oracle `o` and a copy of its whole state."""
    g = sg.FlipFluidGPU(geom["density"], geom["width"], geom["height"], geom["spacing"], geom["radius"], geom["max_particles"])
    for k in ("f_num_x", "f_num_y", "p_num_x", "p_num_y"):
        assert getattr(g, k) == getattr(o, k), k
    assert g.h == o.h and g.p_inv_spacing == o.p_inv_spacing and g.f_inv_spacing == o.f_inv_spacing
    push_state(o, g)
    return g


def push_state(o, g, particles=True, grid=True):
    if particles:
        g.num_particles = o.num_particles
        P = o.num_particles
        g.upload("particle_pos", o.particle_pos[:2 * P])
        g.upload("particle_vel", o.particle_vel[:2 * P])
        g.upload("particle_color", o.particle_color[:3 * P])
    if grid:
        for n in GRID_FIELDS:
            g.upload(n, getattr(o, n))
        g.upload("cell_type", o.cell_type)
        g.particle_rest_density = float(o.particle_rest_density)


def clone_oracle(o, geom, kind="port"):
    c = co.FlipOracle(geom["density"], geom["width"], geom["height"], geom["spacing"], geom["radius"], geom["max_particles"], kind=kind)
    c.set_num_particles(o.num_particles)
    for n in co._FLIP_FIELDS:
        getattr(c, n)[:] = getattr(o, n)
    c.set_rest_density(float(o.particle_rest_density))
    return c


def bits(a):
    return np.ascontiguousarray(a).view(np.int32)


def assert_bit_equal(a, b, what=""):
    a = np.asarray(a).reshape(-1)
    b = np.asarray(b).reshape(-1)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.dtype.kind == "f":
        same = bits(a.astype(np.float32, copy=False)) == bits(b.astype(np.float32, copy=False))
    else:
        same = a == b
    if not same.all():
        idx = np.flatnonzero(~same)
        raise AssertionError(f"{what}: {idx.size} of {a.size} elements differ, first at {idx[0]}: {a[idx[0]]!r} vs {b[idx[0]]!r}")


def overlap_stats(pos, r):
    """(pairs with d<2r, min d / r, mean penetration / r) -- the T3 acceptance statistics"""
    from scipy.spatial import cKDTree
    p = np.asarray(pos, np.float64).reshape(-1, 2)
    pairs = cKDTree(p).query_pairs(2 * r * 0.999, output_type="ndarray")
    if len(pairs) == 0:
        return 0, 2.0, 0.0
    d = np.linalg.norm(p[pairs[:, 0]] - p[pairs[:, 1]], axis=1)
    return len(pairs), float(d.min() / r), float((2 * r - d).mean() / r)
