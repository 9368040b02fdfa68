"""Per-step phase timings + P2G margin for the synthetic dam-break (diagnostic, GPU box)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import simsandgames_b200 as sg
N = int(sys.argv[1]); K = int(sys.argv[2])
sim, kw, g = sg.scenes.make_flip_tank("synthetic", N=N)
sim.enable_timings(True)
for k in range(K):
    sim.reset_timings()
    sim.simulate(**kw, n_steps=1, update_colors=False)
    sim.synchronize()
    t = sim.timings()
    cnt = sim.readback("num_cell_particles") if (k % 5 == 4 or k == K - 1) else None
    print(k, {a: round(b, 2) for a, b in t.items() if b > 0.01},
          "" if cnt is None else f"max/cell {cnt.max()} nonempty {(cnt>0).sum()}", flush=True)
