"""Stability probe: max |vel|, particles per hash cell, P2G margin per step (diagnostic, GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import simsandgames_b200 as sg
N = int(sys.argv[1]); K = int(sys.argv[2]); order = int(sys.argv[3]) if len(sys.argv) > 3 else 0
sep = int(sys.argv[4]) if len(sys.argv) > 4 else 1
sim, kw, g = sg.scenes.make_flip_tank("synthetic", N=N)
kw = dict(kw, separate_particles=bool(sep))
for k in range(K):
    sim.simulate(**kw, n_steps=1, update_colors=False, solver_order=order)
    if k % 3 == 2 or k == K - 1:
        v = sim.readback("particle_vel"); cnt = sim.readback("num_cell_particles")
        print(k, "vmax", np.abs(v).max(), "cells/step", np.abs(v).max() * kw["dt"] / float(sim.h), "max/cell", cnt.max(),
              "umax", np.abs(sim.readback("u")).max(), "pmax", np.abs(sim.readback("p")).max(), flush=True)
