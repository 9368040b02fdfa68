"""GPU: the C++ headless driver (reference demo loop, rendering stripped) on the C++ facade FlipFluidGPU/FluidGPU,
against the same driver source compiled with the unmodified reference header (oracle/_ref/headless_ref)."""
import json
import os
import subprocess
import numpy as np
import pytest
from oracle import cpu_oracle as co

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU_BIN = os.path.join(ROOT, "simsandgames_b200", "host", "headless_gpu")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "headless_ref")


def run(binary, *args):
    out = subprocess.run([binary, *map(str, args)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(GPU_BIN):
        subprocess.run(["make", "-C", os.path.dirname(GPU_BIN)], check=True)


def test_flip_default_tank_through_cpp_facade():
    g = run(GPU_BIN, "flip", "default", 5)
    assert g["grid"] == [134, 101] and g["particles"] == 19092 and g["impl"] == "flipgpu"
    o, p, _ = co.make_flip_scene("default", oracle_kind="port")
    o.simulate(**p, n_steps=5)
    P = o.num_particles
    # 5 steps from the lattice: the separation schedule and the P2G add order differ (measured 1.1e-3 on mean_x, width 4)
    assert abs(g["mean_x"] - float(o.particle_pos[0:2 * P:2].astype(np.float64).mean())) < 5e-3
    assert abs(g["mean_y"] - float(o.particle_pos[1:2 * P:2].astype(np.float64).mean())) < 5e-3
    assert abs(g["cell_color_sum"] - float(o.cell_color.astype(np.float64).sum())) < 1e-2 * float(o.cell_color.sum())
    if os.path.exists(REF_BIN):
        r = run(REF_BIN, "flip", "default", 5)
        assert r["impl"] == "reference" and r["particles"] == g["particles"] and r["grid"] == g["grid"]
        assert abs(r["mean_y"] - g["mean_y"]) < 5e-3 and abs(r["mean_x"] - g["mean_x"]) < 5e-3


def test_flip_long_run_invariant_through_cpp_facade():
    g = run(GPU_BIN, "flip", "default", 300)
    o, p, _ = co.make_flip_scene("default", oracle_kind="port")
    o.simulate(**p, n_steps=300)
    P = o.num_particles
    assert abs(g["mean_y"] - float(o.particle_pos[1:2 * P:2].astype(np.float64).mean())) < 3e-2 * 3.0


def test_euler_wind_tunnel_through_cpp_facade():
    g = run(GPU_BIN, "euler", 96, 72, 40, 30)
    e = co.make_euler_scene(96, 72, oracle_kind="port")
    e.step(40, 1 / 60.0, n_steps=30)
    assert g["grid"] == [98, 74]
    # default order = lexicographic wavefront: bit-identical fields, so the means agree to summation noise
    assert abs(g["smoke_mean"] - float(e.m.astype(np.float64).mean())) < 1e-9
    assert abs(g["abs_pressure_mean"] - float(np.abs(e.pressure).astype(np.float64).mean())) < 1e-6 * float(np.abs(e.pressure).mean()) + 1e-9


def test_frame_dump_identical_to_reference_frame(tmp_path):
    """the frame the UI would draw (cell colours + particles), written by the same driver source built against
    the CUDA facade (reference-order mode) and against the unmodified reference header: the two PPM files are identical."""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/headless_ref not built here")
    a, b = str(tmp_path / "gpu.ppm"), str(tmp_path / "ref.ppm")
    env = dict(os.environ, FLIPGPU_EXACT="1")
    out = subprocess.run([GPU_BIN, "flip", "default", "40", "--ppm", a], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr
    out = subprocess.run([REF_BIN, "flip", "default", "40", "--ppm", b], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    ga, gb = open(a, "rb").read(), open(b, "rb").read()
    assert len(ga) == len(gb) > 500000 and ga == gb


def test_euler_frame_dump(tmp_path):
    path = str(tmp_path / "smoke.ppm")
    out = subprocess.run([GPU_BIN, "euler", "96", "72", "40", "30", "--ppm", path], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    data = open(path, "rb").read()
    assert data.startswith(b"P6\n98 74\n255\n")
    e = co.make_euler_scene(96, 72, oracle_kind="port")
    e.step(40, 1 / 60.0, n_steps=30)
    img = np.frombuffer(data[len(b"P6\n98 74\n255\n"):], np.uint8).reshape(74, 98, 3)[..., 0]
    want = np.where(e.solid.T != 0, 0, np.clip(e.m.T, 0, 1) * 255 + .5).astype(np.uint8)
    assert np.array_equal(img, want)
