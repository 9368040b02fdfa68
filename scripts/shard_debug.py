"""2-rank debug of the sharded FLIP step: per-step particle accounting (run under torch.distributed.run)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simsandgames_b200 as sg
from simsandgames_b200 import parallel
from oracle import cpu_oracle as co
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
uid = parallel.broadcast_unique_id()
o, p, g = co.make_flip_scene("default", oracle_kind="port")
o.simulate(**p, n_steps=40)
P = o.num_particles
pos, vel = o.particle_pos[:2 * P].reshape(-1, 2), o.particle_vel[:2 * P].reshape(-1, 2)
sim = sg.FlipFluidSharded(g["density"], g["width"], g["height"], g["spacing"], g["radius"], 2 * P, P, rank, world, unique_id=uid, device=local)
for name in ("s", "u", "v"):
    sim.upload(name, getattr(o, name))
sim.particle_rest_density = float(o.particle_rest_density)
mine = sim.owns(pos)
sim.upload_particles(pos[mine], vel[mine], np.flatnonzero(mine).astype(np.int32))
kw = dict(dt=p["dt"], gravity=9.81, flip_ratio=0.9, num_pressure_iters=50, num_particle_iters=2, over_relaxation=1.9, compensate_drift=True,
          separate_particles=True, obstacle_x=p["ox"], obstacle_y=p["oy"], obstacle_vel_x=0.0, obstacle_vel_y=0.0, obstacle_radius=p["orad"])
prev = {}
for step in range(10):
    sim.simulate(**kw, n_steps=1)
    gp, gv, gi = sim.readback_particles()
    allp = [None] * world
    dist.all_gather_object(allp, (gp, gv, gi))
    ids = np.concatenate([a[2] for a in allp])
    if rank == 0:
        uniq, cnt = np.unique(ids, return_counts=True)
        lost = np.setdiff1d(np.arange(P), uniq)
        print(f"step {step}: owned {[len(a[2]) for a in allp]} total {len(ids)} dup {int((cnt > 1).sum())} lost {len(lost)}", flush=True)
        if len(lost) and prev:
            h = float(sim.h)
            for i in lost[:6]:
                pp, vv, rr = prev[i]
                print(f"   lost id {i}: prev pos {pp} row {pp[1] / h:.2f} vel {vv} (cells/step {vv * kw['dt'] / h}) prev owner {rr}; slab boundary row {sim.j1 if rank == 0 else sim.j0}", flush=True)
        prev = {}
        for r, a in enumerate(allp):
            for k, i in enumerate(a[2]):
                prev[int(i)] = (a[0][2 * k:2 * k + 2], a[1][2 * k:2 * k + 2], r)
dist.barrier(); dist.destroy_process_group()
