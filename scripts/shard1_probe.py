"""One rank's slab-sharded step WITHOUT peers (nranks = 1: same kernels as a rank of the multi-GPU run, no NCCL), S2 tank: CUDA-event
phases, and one step between cudaProfilerStart/Stop for `ncu --profile-from-start off` (diagnostic, GPU box).
usage: python scripts/shard1_probe.py [N=4096]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import simsandgames_b200 as sg
from simsandgames_b200 import scenes
f32 = np.float32
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sp = f32(3) / f32(N)
width = float((f32(N) - f32(.5)) * sp); height = width
radius = float(f32(.19) * sp)
dt = float((f32(1) / f32(60)) * (f32(100) / f32(N)))
pos, _, P0 = scenes.tank_particles(width, height, float(sp), radius, want_s=False)
sim = sg.FlipFluidSharded(1000.0, width, height, float(sp), radius, int(P0 * 1.3), P0, 0, 1, device=0)
s = np.ones((N, N), np.float32); s[:, 0] = 0; s[:, -1] = 0; s[0] = 0; s[-1] = 0
h_s = torch.from_numpy(s.reshape(-1)).pin_memory()
sim.upload_rows_from("s", h_s.data_ptr(), 0, N)
sim.upload_particles(pos, None, np.arange(P0, dtype=np.int32))
kw = dict(dt=dt, gravity=9.81, flip_ratio=0.9, num_pressure_iters=50, num_particle_iters=2, over_relaxation=1.9, compensate_drift=True,
          separate_particles=True, obstacle_x=0.0, obstacle_y=0.0, obstacle_vel_x=0.0, obstacle_vel_y=0.0, obstacle_radius=0.0)
sim.simulate(**kw, n_steps=4)
sim.synchronize()
sim.enable_timings(True); sim.reset_timings()
K = 5
sim.simulate(**kw, n_steps=K)
sim.synchronize()
print(json.dumps({"N": N, "P": P0, "ms": {k: round(v / K, 3) for k, v in sim.timings().items()}, "total_ms": round(sum(sim.timings().values()) / K, 3)}), flush=True)
sim.enable_timings(False)
torch.cuda.profiler.start()
sim.simulate(**kw, n_steps=1)
sim.synchronize()
torch.cuda.profiler.stop()
sim.close()
