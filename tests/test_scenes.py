"""CPU: the product's scene builders (host C++ in libflipgpu.so + simsandgames_b200/scenes.py) against the
oracle's independent numpy restatement of the reference UI shells, and against the SURVEY's particle counts."""
import numpy as np
import pytest
from oracle import cpu_oracle as co
from simsandgames_b200 import scenes


@pytest.mark.parametrize("kind,N", [("default", None), ("synthetic", 64), ("synthetic", 256)])
def test_tank_lattice_bit_equal(kind, N):
    g = scenes.tank_geometry(kind, N)
    o, p, og = co.make_flip_scene(kind, N=N, oracle_kind="port")
    for k in ("width", "height", "spacing", "radius"):
        assert np.float32(g[k]) == np.float32(og[k]), k
    pos, s, P = scenes.tank_particles(g["width"], g["height"], g["spacing"], g["radius"], o.f_num_x, o.f_num_y)
    assert P == o.num_particles
    assert np.array_equal(pos.view(np.int32), o.particle_pos[:2 * P].view(np.int32))
    walls = np.ones((o.f_num_y, o.f_num_x), np.float32); walls[0] = 0; walls[-1] = 0; walls[:, 0] = 0; walls[:, -1] = 0
    assert np.array_equal(s.reshape(walls.shape), walls)
    assert abs(g["dt"] - p["dt"]) == 0 and g["iters"] == p["num_pressure_iters"]


@pytest.mark.parametrize("N,P", [(1024, 3990320), (4096, 64264080), (8192, 257311935), (16384, 1029824840)])
def test_particle_counts_of_the_baseline_configs(N, P):
    """S2 = 64 264 080, S4 = 257 311 935, S3 = 1 029 824 840 particles."""
    assert scenes._count_only(scenes.tank_geometry("synthetic", N)) == P


def test_wind_tunnel_arrays_match_oracle_scene():
    for x, y in ((96, 72), (40, 33)):
        o = co.make_euler_scene(x, y, oracle_kind="port")
        u, m, solid = scenes.wind_tunnel_arrays(o.num_x, o.num_y)
        assert np.array_equal(u, o.u) and np.array_equal(m, o.m) and np.array_equal(solid, o.solid)


@pytest.mark.parametrize("N,world", [(256, 2), (256, 4), (512, 8)])
def test_per_rank_lattice_of_the_sharded_bench_is_the_global_lattice(N, world):
    """bench.py builds each rank's particles without materialising the others (needed at 16384^2 / 1.03e9 particles):
    the union over ranks is exactly the global lattice of flip_scene_tank, ids included, and every particle lands on its owner."""
    import bench
    from simsandgames_b200 import parallel
    g = scenes.tank_geometry("synthetic", N)
    pos, _, P = scenes.tank_particles(g["width"], g["height"], g["spacing"], g["radius"], want_s=False)
    pos = pos.reshape(-1, 2)
    seen = np.zeros(P, bool)
    for rank, (j0, j1) in enumerate(parallel.slab_partition(N, world)):
        lat = bench.synthetic_rank_lattice(N, j0, j1)
        assert lat["P_global"] == P and lat["dt"] == g["dt"] and lat["n_own"] == lat["ids"].size
        assert not seen[lat["ids"]].any()
        seen[lat["ids"]] = True
        assert np.array_equal(lat["pos"].view(np.int32), pos[lat["ids"]].view(np.int32))
        row = np.clip((np.float32(1) / lat["h_grid"] * lat["pos"][:, 1]).astype(np.int32), 0, N - 1)
        assert ((row >= j0) & (row < j1)).all()
    assert seen.all()
