"""Worker of the multi-rank tests (launched with torch.distributed.run, one process per rank).
usage: multi_worker.py <inputs.npz> <out_prefix> <backend gloo|nccl>
GPU mode (a CUDA device per rank): runs this rank's slab of the sharded red-black projection and saves its rows.
CPU mode (no CUDA): only exercises the host logic -- id broadcast, partition, gather."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simsandgames_b200 import parallel  # noqa: E402


def run_flip(d, out_prefix, rank, world, local, uid, sg):
    """this rank's slab of a sharded FLIP run: global scene in, owned particles + own grid rows out"""
    g = {k: float(d[k]) for k in ("density", "width", "height", "spacing", "radius")}
    pos, vel = d["pos"].reshape(-1, 2), d["vel"].reshape(-1, 2)
    rows = [int(x) for x in d["row_starts"]] if "row_starts" in d.files else None     # caller-chosen (unequal) slabs
    sim = sg.FlipFluidSharded(g["density"], g["width"], g["height"], g["spacing"], g["radius"], int(d["capacity"]), pos.shape[0], rank, world,
                              unique_id=uid, device=local, row_starts=rows)
    for name in ("s", "u", "v"):
        sim.upload(name, d[name])                      # grid arrays keep the global shape on every rank
    if float(d["rest"]) != 0.0:
        sim.particle_rest_density = float(d["rest"])
    mine = sim.owns(pos)
    ids = np.flatnonzero(mine).astype(np.int32)
    sim.upload_particles(pos[mine], vel[mine], ids)
    kw = dict(dt=float(d["dt"]), gravity=9.81, flip_ratio=0.9, num_pressure_iters=int(d["iters"]), num_particle_iters=2,
              over_relaxation=1.9, compensate_drift=True, separate_particles=True, obstacle_x=float(d["ox"]), obstacle_y=float(d["oy"]),
              obstacle_vel_x=0.0, obstacle_vel_y=0.0, obstacle_radius=float(d["orad"]))
    sim.simulate(**kw, n_steps=int(d["flip_steps"]))
    p, v, i = sim.readback_particles()
    np.savez(f"{out_prefix}_rank{rank}.npz", pos=p, vel=v, ids=i, u=sim.readback_own_rows("u"), v=sim.readback_own_rows("v"),
             p=sim.readback_own_rows("p"), cell_type=sim.readback_own_rows("cell_type"), density=sim.readback_own_rows("particle_density"),
             rest=np.float32(sim.particle_rest_density), info=np.array(sim.shard_info()))
    sim.close()


def main():
    inputs, out_prefix, backend = sys.argv[1], sys.argv[2], sys.argv[3]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    has_gpu = torch.cuda.is_available() and torch.cuda.device_count() > local
    if has_gpu:
        torch.cuda.set_device(local)
    dist.init_process_group(backend, rank=rank, world_size=world)
    d = np.load(inputs)
    ns, nf = d["vel_fast"].shape if "vel_fast" in d.files else (0, 0)
    part = parallel.slab_partition(ns, world)
    if not has_gpu:
        uid = parallel.broadcast_unique_id(lambda: bytes(range(128)))
        assert uid == bytes(range(128))
        j0, j1 = part[rank]
        own = d["vel_fast"][j0:j1] * 2.0
        gathered = [None] * world
        dist.all_gather_object(gathered, own)
        if rank == 0:
            np.save(out_prefix + "_stitched.npy", parallel.stitch_rows(gathered))
        dist.barrier()
        dist.destroy_process_group()
        return
    import simsandgames_b200 as sg
    uid = parallel.broadcast_unique_id()
    if "flip_steps" in d.files:
        run_flip(d, out_prefix, rank, world, local, uid, sg)
        dist.barrier()
        dist.destroy_process_group()
        return
    sh = sg.ShardedProjection(nf, ns, bool(d["x_fast"]), rank, world, unique_id=uid, device=local)
    assert (sh.j0, sh.j1) == part[rank]
    for name in ("vel_fast", "vel_slow", "p", "s", "density", "cell_type"):
        sh.upload_window(name, d[name])
    sh.solve(int(d["iters"]), float(d["cp"]), float(d["omega"]), bool(d["drift"]), float(d["rest"]))
    sh.synchronize()
    np.savez(f"{out_prefix}_rank{rank}.npz", vel_fast=sh.readback_own("vel_fast"), vel_slow=sh.readback_own("vel_slow"),
             p=sh.readback_own("p"), counters=np.array(sh.counters()))
    dist.barrier()
    sh.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
