"""2-rank debug of the sharded FLIP step against the one-GPU run, field by field after each step."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simsandgames_b200 as sg
from simsandgames_b200 import parallel
from oracle import cpu_oracle as co
from tests.helpers import gpu_like, product_args
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
uid = parallel.broadcast_unique_id()
o, p, g = co.make_flip_scene("default", oracle_kind="port")
o.simulate(**p, n_steps=int(sys.argv[1]) if len(sys.argv) > 1 else 40)
P = o.num_particles
pos, vel = o.particle_pos[:2 * P].reshape(-1, 2), o.particle_vel[:2 * P].reshape(-1, 2)
sim = sg.FlipFluidSharded(g["density"], g["width"], g["height"], g["spacing"], g["radius"], 2 * P, P, rank, world, unique_id=uid, device=local)
for name in ("s", "u", "v"):
    sim.upload(name, getattr(o, name))
sim.particle_rest_density = float(o.particle_rest_density)
mine = sim.owns(pos)
sim.upload_particles(pos[mine], vel[mine], np.flatnonzero(mine).astype(np.int32))
kw = product_args(p)
one = gpu_like(o, g) if rank == 0 else None
ny, nx = o.f_num_y, o.f_num_x
for step in range(3):
    sim.simulate(**kw, n_steps=1)
    gp, gv, gi = sim.readback_particles()
    rows = {k: sim.readback_own_rows(n) for k, n in (("u", "u"), ("v", "v"), ("p", "p"), ("cell_type", "cell_type"), ("density", "particle_density"), ("prev_u", "prev_u"))}
    allp = [None] * world
    dist.all_gather_object(allp, (gp, gv, gi, rows, sim.shard_info()))
    if rank == 0:
        one.simulate(**kw, n_steps=1, solver_order=sg.RED_BLACK, update_colors=False)
        ids = np.concatenate([a[2] for a in allp])
        gpos = np.empty((P, 2), np.float32); gpos[ids] = np.concatenate([a[0] for a in allp]).reshape(-1, 2)
        gvel = np.empty((P, 2), np.float32); gvel[ids] = np.concatenate([a[1] for a in allp]).reshape(-1, 2)
        wp, wv = one.readback("particle_pos").reshape(-1, 2), one.readback("particle_vel").reshape(-1, 2)
        bad = np.flatnonzero((gpos != wp).any(1))
        print(f"step {step}: info {[a[4] for a in allp]} pos differ {len(bad)} max {np.abs(gpos - wp).max():.3e}; vel differ {int((gvel != wv).any(1).sum())} max {np.abs(gvel - wv).max():.3e}", flush=True)
        if len(bad):
            h = float(sim.h)
            print("   rows of differing particles:", np.unique((wp[bad, 1] / h).astype(int))[:40], flush=True)
        for k, name in (("u", "u"), ("v", "v"), ("p", "p"), ("cell_type", "cell_type"), ("density", "particle_density"), ("prev_u", "prev_u")):
            got = np.concatenate([a[3][k] for a in allp], axis=0)
            want = one.readback(name).reshape(ny, nx)
            d = got != want
            print(f"   {k:10s} differ {int(d.sum()):6d} rows {np.unique(np.nonzero(d)[0])[:30]}", flush=True)
dist.barrier(); dist.destroy_process_group()
