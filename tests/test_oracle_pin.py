"""CPU: the C port (oracle/flip_oracle.c, euler_oracle.c) is pinned to the unmodified reference --
(a) against the committed golden hashes made from oracle/_ref by tests/golden/make_golden.py,
(b) live, bit for bit, against oracle/_ref when that library is present (build container)."""
import json
import os
import numpy as np
import pytest
from oracle import cpu_oracle as co
from tests.golden.make_golden import flip_record, sha

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.json")))


@pytest.mark.parametrize("case,kind,N", [("flip_default", "default", None), ("flip_synthetic_64", "synthetic", 64)])
def test_flip_port_matches_reference_golden(case, kind, N):
    g = GOLD[case]
    o, p, geom = co.make_flip_scene(kind, N=N, oracle_kind="port")
    assert dict(f_num_x=o.f_num_x, f_num_y=o.f_num_y, p_num_x=o.p_num_x, p_num_y=o.p_num_y, num_particles=o.num_particles,
                h_bits=int(np.float32(o.h).view(np.int32)), p_inv_bits=int(np.float32(o.p_inv_spacing).view(np.int32))) == g["dims"]
    assert sha(o.particle_pos) == g["scene"]["particle_pos"] and sha(o.s) == g["scene"]["s"]
    done = 0
    for c in sorted(int(k) for k in g["steps"]):
        o.simulate(**p, n_steps=c - done); done = c
        rec = flip_record(o)
        for k, v in g["steps"][str(c)].items():
            assert rec[k] == v, (case, c, k)


def test_default_tank_numbers_from_survey():
    """134x101 cells, 19 092 particles, hash grid 203x152, latched rest density 3.130026."""
    d = GOLD["flip_default"]["dims"]
    assert (d["f_num_x"], d["f_num_y"], d["num_particles"], d["p_num_x"], d["p_num_y"]) == (134, 101, 19092, 203, 152)
    rest = np.int32(GOLD["flip_default"]["steps"]["20"]["rest_density_bits"]).view(np.float32)
    assert abs(float(rest) - 3.130026) < 2e-6


@pytest.mark.parametrize("case,x,y", [("euler_96x72", 96, 72), ("euler_33x20", 33, 20)])
def test_euler_port_matches_reference_golden(case, x, y):
    g = GOLD[case]
    o = co.make_euler_scene(x, y, oracle_kind="port")
    assert dict(num_x=o.num_x, num_y=o.num_y) == g["dims"]
    for n in ("u", "m", "solid"):
        assert sha(getattr(o, n)) == g["scene"][n]
    done = 0
    for c in sorted(int(k) for k in g["steps"]):
        o.step(g["iters"], 1 / 60.0, n_steps=c - done); done = c
        for n in ("u", "v", "m", "pressure"):
            assert sha(getattr(o, n)) == g["steps"][str(c)][n], (case, c, n)


def test_port_equals_reference_live(has_ref):
    if not has_ref:
        pytest.skip("oracle/_ref not built here (no /root/reference)")
    a, p, g = co.make_flip_scene("default", oracle_kind="ref")
    b, _, _ = co.make_flip_scene("default", oracle_kind="port")
    for _ in range(3):
        a.simulate(**p, n_steps=15); b.simulate(**p, n_steps=15)
        for n in co._FLIP_FIELDS:
            x, y = getattr(a, n), getattr(b, n)
            assert np.array_equal(x.view(np.int32), y.view(np.int32)), n
    # phase-by-phase entry points agree too
    for fn, args in (("integrate", (1 / 60, 9.81)), ("push_apart", (2,)), ("collide", (3.0, 2.0, 0.1, -0.1, 0.15)),
                     ("transfer", (True,)), ("update_density", ()), ("solve", (50, 1 / 60, 1.9, True)), ("transfer", (False, 0.9)),
                     ("particle_colors", ()), ("cell_colors", ())):
        getattr(a, fn)(*args); getattr(b, fn)(*args)
    for n in co._FLIP_FIELDS:
        assert np.array_equal(getattr(a, n).view(np.int32), getattr(b, n).view(np.int32)), n
    e1, e2 = co.make_euler_scene(oracle_kind="ref"), co.make_euler_scene(oracle_kind="port")
    e1.step(40, 1 / 60, n_steps=25); e2.step(40, 1 / 60, n_steps=25)
    for n in ("u", "v", "m", "pressure"):
        assert np.array_equal(getattr(e1, n), getattr(e2, n)), n
    assert e1.sample(0.31, 0.42, 2) == e2.sample(0.31, 0.42, 2)


def test_binning_invariants_of_the_oracle():
    """flip_fluid.h:169-198: exclusive starts, guard entry, descending ids inside a cell (Q7)."""
    o, p, g = co.make_flip_scene("default", oracle_kind="port")
    o.simulate(**p, n_steps=30)
    P = o.num_particles
    first, cnt, ids = o.first_cell_particle, o.num_cell_particles, o.cell_particle_ids[:P]
    assert first[0] == 0 and first[-1] == P and np.array_equal(np.diff(first), cnt)
    assert np.array_equal(np.sort(ids), np.arange(P))
    for c in np.flatnonzero(cnt > 1)[:500]:
        seg = ids[first[c]:first[c + 1]]
        assert (np.diff(seg) < 0).all()


@pytest.mark.parametrize("kind", ["port", "ref"])
def test_euler_set_obstacle_restatements_agree(kind, has_ref):
    """FluidUI::setObstacle (fluid_ui.h:24-62) cannot be compiled from the reference (its header needs the windowing
    engine): the C restatements (port, and the harness acting on the reference's own Fluid object) must equal the
    independent numpy one, incl. a disc that pokes out of the grid and a dragged obstacle carrying velocity."""
    if kind == "ref" and not has_ref:
        pytest.skip("oracle/_ref not built")
    for (x, y, r, vx, vy) in [(34, 37, 12, 0.0, 0.0), (3, 70, 9, 1.5, -0.25), (95, 2, 20, -3.0, 2.0), (50, 30, 0, 1.0, 1.0)]:
        a = co.EulerOracle(96, 72, 1000.0, 1 / 100.0, kind=kind)
        b = co.EulerOracle(96, 72, 1000.0, 1 / 100.0, kind="port")
        rng = np.random.default_rng(5)
        for o in (a, b):
            o.u[:] = rng.standard_normal(o.u.shape).astype(np.float32) if o is a else a.u
            o.v[:] = 0.5
            o.m[:] = 0.25
            o.solid[:] = 1          # stale obstacle everywhere: the call must clear it
        a.set_obstacle(x, y, r, vx, vy)
        co.euler_set_obstacle(b, x, y, r, np.float32(vx), np.float32(vy))
        for n in ("solid", "m", "u", "v"):
            assert np.array_equal(getattr(a, n), getattr(b, n)), (kind, n, x, y, r)


def test_fluid3d_port_pinned_to_reference_and_goldens(has_ref):
    """eulerian_fluid/src/fluid3d.h: the C port equals the reference header run here (when oracle/_ref is built) and the
    committed hashes of the reference's own output -- project / extrapolate / advectVel (incl. its k < num_y bound) / advectSmoke."""
    import hashlib, json, os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.json")))
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    for key, (x, y, z) in (("fluid3d_30x22x26", (30, 22, 26)), ("fluid3d_12x9x17", (12, 9, 17))):
        rec = gold[key]
        o = co.make_fluid3d_scene(x, y, z, oracle_kind="port")
        assert (o.num_x, o.num_y, o.num_z) == (rec["dims"]["num_x"], rec["dims"]["num_y"], rec["dims"]["num_z"])
        for n, hsh in rec["scene"].items():
            assert sha(getattr(o, n)) == hsh, (key, "scene", n)
        done = 0
        for c in sorted(rec["steps"], key=int):
            o.step(rec["iters"], 1 / 60.0, n_steps=int(c) - done); done = int(c)
            for n in ("u", "v", "w", "m", "pressure"):
                assert sha(getattr(o, n)) == rec["steps"][c][n], (key, c, n)
    if has_ref and co.available("fluid3d", "ref"):
        a, b = co.make_fluid3d_scene(oracle_kind="ref"), co.make_fluid3d_scene(oracle_kind="port")
        for fn, args in (("solve", (15, 1 / 60)), ("extrapolate", ()), ("advect_vel", (1 / 60,)), ("advect_smoke", (1 / 60,)), ("step", (9, 1 / 60, 4))):
            getattr(a, fn)(*args); getattr(b, fn)(*args)
            for n in ("u", "v", "w", "m", "pressure", "new_u", "new_m"):
                assert np.array_equal(getattr(a, n).view(np.int32), getattr(b, n).view(np.int32)), (fn, n)
        for q in [(0.21, 0.2, 0.33), (0.0, 5.0, 0.1), (0.55, 0.02, 0.61)]:
            for fld in range(4):
                assert a.sample(*q, fld) == b.sample(*q, fld)
