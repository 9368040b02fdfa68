"""CPU: the host model of the register-systolic wavefront schedule (tests/systolic_model.cpp -- the dataflow of
simsandgames_b200/csrc/sor_systolic.cu with the warp's lanes as array elements) reproduces the plain lexicographic
Gauss-Seidel loops of flip_fluid.h:458-494 bit for bit: strips skewed one column per stage, rotate shuffles, the
hand-over at the wrap lane, chunk-to-chunk dependencies, open border cells, every neighbour-openness code."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_systolic_schedule_model_is_the_lexicographic_order(tmp_path):
    exe = str(tmp_path / "systolic_model")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "systolic_model.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-3000:]
