"""BASELINE configs[4] on ONE GPU: synthetic 8192^2 tank with the static disc obstacle, 200 pressure sweeps (diagnostic, GPU box)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simsandgames_b200 as sg
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
sim, kw, g = sg.scenes.make_flip_tank("synthetic", N=N, with_obstacle=True)
sim.simulate(**kw, n_steps=2, update_colors=False)
sim.synchronize(); sim.enable_timings(True); sim.reset_timings()
K = 4
sim.simulate(**kw, n_steps=K, update_colors=False); sim.synchronize()
t = sim.timings(); tot = sum(t.values()) / K
interior = (sim.f_num_x - 2) * (sim.f_num_y - 2)
print(json.dumps({"config": f"flip_fluid synthetic {N}^2 with obstacle, {kw['num_pressure_iters']} SOR iters (BASELINE configs[4]), 1 GPU",
                  "particles": sim.num_particles, "ms_per_step": round(tot, 2), "particle_steps_per_s": sim.num_particles / (tot * 1e-3),
                  "cell_updates_per_s": interior * kw["num_pressure_iters"] / (t["solve"] / K * 1e-3),
                  "phase_ms": {k: round(v / K, 2) for k, v in t.items()}, "residual_max_rms": sim.residual()}))
