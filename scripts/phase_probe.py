"""Per-phase CUDA-event timings of flip_step for a synthetic dam-break of N^2 cells (diagnostic, GPU box), plus a hash of
the final particle state so that runs with different kernel variants (FLIPGPU_SOR_V2=1, FLIPGPU_RB_V2=1, ... -- switches
that are read once per process) can be compared for bit-identity across processes.
usage: [ORDER=lex|rb] [SWITCH=1 ...] python scripts/phase_probe.py N [N ...]"""
import sys, os, json, hashlib
ORDER_NAME = os.environ.get('ORDER', 'lex')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simsandgames_b200 as sg
ORDER = sg.RED_BLACK if ORDER_NAME in ('rb', 'red_black') else sg.LEX_WAVEFRONT
switches = {k: v for k, v in os.environ.items() if k.startswith("FLIPGPU_")}
for N in [int(a) for a in sys.argv[1:]] or [1024]:
    sim, kw, g = sg.scenes.make_flip_tank("synthetic", N=N)
    sim.simulate(**kw, n_steps=3, update_colors=False, solver_order=ORDER)
    sim.synchronize()
    sim.enable_timings(True); sim.reset_timings()
    K = 5
    sim.simulate(**kw, n_steps=K, update_colors=False, solver_order=ORDER)
    sim.synchronize()
    t = sim.timings()
    h = hashlib.sha256()
    for name in ("particle_pos", "particle_vel", "u", "v", "p"):
        h.update(sim.readback(name).tobytes())
    print(json.dumps({"N": N, "P": sim.num_particles, "order": ORDER_NAME, "switches": switches,
                      "ms": {k: round(v / K, 3) for k, v in t.items()}, "total_ms": round(sum(t.values()) / K, 3),
                      "state_sha256": h.hexdigest()[:16]}), flush=True)
    sim.close()
