"""A/B per-phase CUDA-event timings of flip_step on one scene (diagnostic, GPU box): the kernel variants selected by the
FLIPGPU_*_V1 environment switches are read at every call, so both run in one process on the same sim.
usage: python scripts/ab_probe.py [N=4096] [K=8]   ->  one JSON line per variant + a bit-equality check of the results."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simsandgames_b200 as sg

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
SWITCHES = ("FLIPGPU_P2G_V1",)   # every *_V1 switch the library currently reads
sim, kw, g = sg.scenes.make_flip_tank("synthetic", N=N)
pos0 = sim.readback("particle_pos").copy()
zeros_v = np.zeros_like(pos0)
zero_c = np.zeros(sim.f_num_cells, np.float32)
results = {}
for label, val in (("new", "0"), ("v1", "1"), ("new_again", "0")):
    for s in SWITCHES:
        os.environ[s] = val
    sim.upload("particle_pos", pos0); sim.upload("particle_vel", zeros_v)
    for n in ("u", "v", "prev_u", "prev_v", "p", "particle_density"):
        sim.upload(n, zero_c)
    sim.particle_rest_density = 0.0
    sim.simulate(**kw, n_steps=3, update_colors=False)
    sim.synchronize()
    sim.enable_timings(True); sim.reset_timings()
    sim.simulate(**kw, n_steps=K, update_colors=False)
    sim.synchronize()
    t = sim.timings()
    sim.enable_timings(False)
    results[label] = (sim.readback("particle_pos").copy(), sim.readback("particle_vel").copy())
    print(json.dumps({"variant": label, "N": N, "P": sim.num_particles, "ms": {k: round(v / K, 3) for k, v in t.items()},
                      "total_ms": round(sum(t.values()) / K, 3)}), flush=True)
same = all(np.array_equal(results["new"][i].view(np.uint32), results["v1"][i].view(np.uint32)) for i in (0, 1))
stable = all(np.array_equal(results["new"][i].view(np.uint32), results["new_again"][i].view(np.uint32)) for i in (0, 1))
print(json.dumps({"bit_identical_new_vs_v1": bool(same), "run_to_run_bit_stable": bool(stable)}), flush=True)
sim.close()
sys.exit(0 if same and stable else 1)
