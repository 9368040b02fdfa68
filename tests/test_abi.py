"""CPU: the C-ABI library loads, exports every symbol include/flipgpu.h declares, and refuses to run without a GPU."""
import ctypes as C
import os
import re
import subprocess
import numpy as np
import pytest
import simsandgames_b200 as sg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "flipgpu.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:flip|euler|flipgpu|fluid3d|sorshard)_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("flip_create", "flip_step", "flip_upload", "flip_readback", "flip_destroy", "flip_set_obstacle",
                 "euler_create", "euler_project", "euler_advect_vel", "euler_advect_smoke", "euler_step", "euler_readback"):
        assert must in names
    assert len(names) >= 40


def test_library_exports_every_declared_symbol():
    lib = sg.load_library()
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.flipgpu_version() >= 100


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "flipgpu.h"\nint main(void){FlipParams p; FlipStepParams s; EulerDims d; (void)p;(void)s;(void)d; return FLIP_FIELD_COUNT+EULER_FIELD_COUNT;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")], check=True)


def test_struct_layouts_match_the_binding():
    """sizeof of the ABI structs as the C compiler sees them == the ctypes mirrors."""
    from simsandgames_b200 import binding as b
    code = '#include <stdio.h>\n#include "flipgpu.h"\nint main(void){printf("%zu %zu %zu %zu\\n", sizeof(FlipParams), sizeof(FlipDims), sizeof(FlipStepParams), sizeof(EulerDims));return 0;}\n'
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(code)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")], check=True)
        out = subprocess.run([os.path.join(d, "s")], capture_output=True, text=True, check=True).stdout.split()
    assert [int(x) for x in out] == [C.sizeof(b._FlipParams), C.sizeof(b._FlipDims), C.sizeof(b._FlipStepParams), C.sizeof(b._EulerDims)]


def test_no_cpu_fallback_without_a_device():
    lib = sg.load_library()
    n = C.c_int(-1)
    rc = lib.flipgpu_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(sg.FlipGpuError, match="no CUDA device"):
        sg.FlipFluidGPU(1000.0, 1.0, 1.0, 0.1, 0.03, 8)
    with pytest.raises(sg.FlipGpuError, match="no CUDA device"):
        sg.FluidGPU(8, 8, 1000.0, 0.01)


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under simsandgames_b200/ or include/ may reference it."""
    bad = []
    for base in ("simsandgames_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                    txt = open(os.path.join(dirpath, f), errors="replace").read()
                    if re.search(r"cpu_oracle|liboracle|libref_|oracle/_|import oracle|from oracle", txt):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
