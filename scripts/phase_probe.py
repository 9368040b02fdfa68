"""Per-phase CUDA-event timings of flip_step for a synthetic dam-break of N^2 cells (diagnostic, GPU box)."""
import sys, os, json
ORDER = int(os.environ.get('ORDER', '1'))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simsandgames_b200 as sg
for N in [int(a) for a in sys.argv[1:]] or [1024]:
    sim, kw, g = sg.scenes.make_flip_tank("synthetic", N=N)
    sim.simulate(**kw, n_steps=3, update_colors=False, solver_order=ORDER)
    sim.synchronize()
    sim.enable_timings(True); sim.reset_timings()
    K = 5
    sim.simulate(**kw, n_steps=K, update_colors=False, solver_order=ORDER)
    sim.synchronize()
    t = sim.timings()
    print(json.dumps({"N": N, "P": sim.num_particles, 
                      "ms": {k: round(v / K, 3) for k, v in t.items()}, "total_ms": round(sum(t.values()) / K, 3)}), flush=True)
    sim.close()
