"""Solver-only timing (diagnostic, GPU box): one pressure solve of the synthetic dam-break state after a few steps, per
solver order, with a bit-comparison of the lexicographic variants against each other.
usage: python scripts/solver_probe.py [N=4096] [iters=50] [with_obstacle=0]"""
import sys, os, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import simsandgames_b200 as sg
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
obst = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
sim, kw, g = sg.scenes.make_flip_tank("synthetic", N=N, with_obstacle=obst)
kw = dict(kw, num_pressure_iters=min(kw["num_pressure_iters"], 50))
sim.simulate(**kw, n_steps=4, update_colors=False)
sim.transferVelocitiesToGrid() if False else None
u0, v0 = sim.readback("u").copy(), sim.readback("v").copy()
stream = torch.cuda.ExternalStream(sim.stream)
res = {}
VARIANTS = (("systolic", sg.LEX_SYSTOLIC), ("tiled", sg.LEX_WAVEFRONT)) if os.environ.get("PROBE_LEX_ONLY") else (("systolic", sg.LEX_SYSTOLIC), ("tiled", sg.LEX_WAVEFRONT), ("red_black", sg.RED_BLACK))
for name, order in VARIANTS:
    ms = []
    for rep in range(4):
        sim.upload("u", u0); sim.upload("v", v0)
        sim.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sim.solveIncompressibility(iters, kw["dt"], 1.9, True, solver_order=order)
        e1.record(stream)
        sim.synchronize(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    res[name] = (min(ms[1:]), sim.readback("u").copy(), sim.readback("v").copy(), sim.readback("p").copy())
    if name == "systolic" and os.environ.get("FLIPGPU_SOR_PROFILE"):
        pr = sim.solver_profile()
        tasks = max(pr[4], 1)
        print(json.dumps({"systolic_profile_cycles_per_task": {"wait": pr[0] / tasks, "flush": pr[1] / tasks, "land_fetch": pr[2] / tasks, "steps": pr[3] / tasks,
                                                              "total": pr[5] / tasks}, "tasks": pr[4]}), flush=True)
same = all(np.array_equal(res["systolic"][i].view(np.int32), res["tiled"][i].view(np.int32)) for i in (1, 2, 3))
print(json.dumps({"N": N, "iters": iters, "obstacle": obst, "ms": {k: round(v[0], 3) for k, v in res.items()},
                  "systolic_equals_tiled_bitwise": bool(same), "env": {k: v for k, v in os.environ.items() if k.startswith("FLIPGPU_")}}), flush=True)
sim.close()
