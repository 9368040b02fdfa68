"""Blocked-solver phase profile (FLIPGPU_SOR_PROFILE=1) on the dam-break (diagnostic, GPU box)."""
import sys, os, ctypes as C
os.environ["FLIPGPU_SOR_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simsandgames_b200 as sg
N = int(sys.argv[1]); K = 5
sim, kw, g = sg.scenes.make_flip_tank("synthetic", N=N)
sim.simulate(**kw, n_steps=3, update_colors=False)
sim.synchronize(); sim.enable_timings(True); sim.reset_timings()
out0 = (C.c_ulonglong * 8)(); sim._lib.flip_debug_solver_profile(sim._h, out0)
sim.simulate(**kw, n_steps=K, update_colors=False); sim.synchronize()
out = (C.c_ulonglong * 8)(); sim._lib.flip_debug_solver_profile(sim._h, out)
d = [out[i] - out0[i] for i in range(5)]
tiles = d[4]
print("solve ms/step", sim.timings()["solve"] / K, "tiles/solve", out[5], "Kc", out[6], "chunks", out[7])
for name, v in zip(("dep wait", "compute", "other", "-"), d[:4]):
    print(f"  {name:10s} {v / tiles:12.0f} cycles/strip  ({v / tiles / 1.9e3:.1f} us @1.9GHz)")
print("  strips per solve", tiles / K)
