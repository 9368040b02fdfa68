"""Profiling target (GPU box): the kernels of one FLIP step at S2 (reference solver order, then red-black) and of one Eulerian
step at S1 (lexicographic, then red-black), launch by launch (graph replay off so every launch is a separate ncu entry).
Only the section between cudaProfilerStart/Stop is captured: run ncu with --profile-from-start off.
usage: ncu --profile-from-start off ... python scripts/ncu_target.py [N=4096]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import simsandgames_b200 as sg
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sim, kw, g = sg.scenes.make_flip_tank("synthetic", N=N)
sim.set_graph_mode(False)
sim.simulate(**kw, n_steps=4, update_colors=False)
sim.simulate(**kw, n_steps=1, update_colors=False, solver_order=sg.RED_BLACK)
sim.synchronize()
torch.cuda.profiler.start()
sim.simulate(**kw, n_steps=1, update_colors=False)                              # step 6: steady state, reference solver order
sim.simulate(**kw, n_steps=1, update_colors=False, solver_order=sg.RED_BLACK)   # + the temporally blocked red-black kernel
sim.synchronize()
torch.cuda.profiler.stop()
sim.close()
e = sg.scenes.make_wind_tunnel(N // 2, N // 2)
e.step(100, 1 / 60.0, n_steps=2, solver_order=sg.LEX_WAVEFRONT)
e.synchronize()
torch.cuda.profiler.start()
e.step(100, 1 / 60.0, n_steps=1, solver_order=sg.LEX_WAVEFRONT)
e.step(100, 1 / 60.0, n_steps=1, solver_order=sg.RED_BLACK)
e.synchronize()
torch.cuda.profiler.stop()
e.close()
print("ncu target done")
