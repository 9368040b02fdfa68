#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native grid-fluid step.

Workload (BASELINE.json configs[2], the config the north_star target is quoted on):
  flip_fluid synthetic dam-break, 4096^2 grid, 64 264 080 particles, simulate(dt, 9.81, .9, 50, 2, 1.9, true, true, ...)
  one "step" = one FlipFluid::simulate (flip_fluid.h:560-581) with the render-state colour kernels off.
Metric: particle-steps/s (whole job).  cell-updates/s of the pressure solve and the roofline are reported beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W]          product arm (CUDA through the C ABI)
  python bench.py --impl reference ...                         the reference's own CPU code on the host cores
Under torchrun (N>1) every rank owns one GPU; timing = CUDA events on the library's stream, bracketed by a
barrier + synchronize, max over ranks.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "flip_particle_steps_per_s"
UNIT = "particle-steps/s"
BENCH_N = 4096
NCU_KERNELS_CSV = os.path.join("profiles", "r2_ncu_kernels.csv")   # per-kernel ncu --set full numbers of THIS workload (scripts/ncu_summary.py)
WORKLOAD = "flip_fluid synthetic dam-break 4096^2 grid, 64264080 particles, 50 SOR iters, 2 separation sweeps (BASELINE configs[2])"


def traffic_from_profile(kernel_prefix):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernel, from the committed ncu table (None if absent)."""
    import csv
    try:
        with open(os.path.join(ROOT, NCU_KERNELS_CSV)) as f:
            for row in csv.DictReader(f):
                if row["kernel"].replace("fg::", "").startswith(kernel_prefix):
                    return float(row["dram_read_bytes"]) + float(row["dram_write_bytes"])
    except Exception:
        pass
    return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        for line in self.f.read().splitlines():
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        return out


def cpu_sample_size(steps, warmup):
    """Grid size of the bounded CPU sample: the same dam-break rule, sized so (steps+warmup) CPU steps take ~2 minutes
    (the serial reference needs ~4.4 s/step at 1024^2 and ~70 s/step at 4096^2)."""
    budget_s = 120.0
    n = 1024.0 * math.sqrt(budget_s / (4.4 * max(1, steps + warmup)))
    return int(max(256, min(1024, (int(n) // 128) * 128)))


def run_cpu_reference(n_sample, steps, warmup):
    """The reference's CPU step (oracle/_ref = unmodified flip_fluid.h; else the C port) on a bounded sample."""
    from oracle import cpu_oracle as co
    kind = co.best_kind("flip")
    o, p, g = co.make_flip_scene("synthetic", N=n_sample, oracle_kind=kind)
    P = o.num_particles
    if warmup:
        o.simulate(**p, n_steps=warmup)
    t0 = time.perf_counter()
    ph = o.simulate(**p, n_steps=steps, timed=True)
    dt = time.perf_counter() - t0
    interior = (o.f_num_x - 2) * (o.f_num_y - 2)
    return dict(kind="reference" if kind == "ref" else "port", particles=P, grid=n_sample, seconds=dt, steps=steps,
                value=P * steps / dt, cell_updates_per_s=interior * p["num_pressure_iters"] * steps / (ph[5] * 1e-9),
                phase_share={k: round(v / sum(ph), 4) for k, v in zip(
                    ("integrate", "pushApart", "collide", "p2g", "density", "solve", "g2p", "colours"), ph)})



def run_cpu_reference_full_size(n=BENCH_N, steps=1):
    """the reference's own CPU step on the FULL benchmark config (4096^2 cells / 64 264 080 particles), one
    step from the initial state (about 80 s on one host core; the reference is strictly serial)."""
    from oracle import cpu_oracle as co
    kind = co.best_kind("flip")
    o, p, g = co.make_flip_scene("synthetic", N=n, oracle_kind=kind)
    P = o.num_particles
    t0 = time.perf_counter()
    ph = o.simulate(**p, n_steps=steps, timed=True)
    dt = time.perf_counter() - t0
    interior = (o.f_num_x - 2) * (o.f_num_y - 2)
    out = dict(kind="reference" if kind == "ref" else "port", grid=[o.f_num_x, o.f_num_y], particles=P, steps=steps, seconds_per_step=dt / steps,
               value=P * steps / dt, unit=UNIT, cores=1, cell_updates_per_s=interior * p["num_pressure_iters"] * steps / (ph[5] * 1e-9),
               phase_share={k: round(v / sum(ph), 4) for k, v in zip(("integrate", "pushApart", "collide", "p2g", "density", "solve", "g2p", "colours"), ph)},
               sample=f"the full BASELINE configs[2] scene, step 1 from the initial lattice, 1 thread")
    o.close() if hasattr(o, "close") else None
    return out


def run_cpu_reference_euler(n_sample=1024, iters=100):
    """eulerian_fluid's CPU step (oracle/_ref = unmodified fluid.h; else the C port) on a bounded sample of BASELINE
    configs[1], timed with and without flush-to-zero/denormals-are-zero: SURVEY Q17 -- the reference stalls on denormals
    while the far field is still quiescent (first steps), so both are reported (8d)."""
    from oracle import cpu_oracle as co
    kind = co.best_kind("euler")
    out = {"kind": "reference" if kind == "ref" else "port", "cores": 1, "unit": "cell-updates/s",
           "sample": f"wind tunnel at {n_sample}^2 interior cells, {iters} pressure iters, steps 1-2 from the initial state, 1 thread"}
    for label, ftz in (("ieee", 0), ("ftz_daz", 1)):
        o = co.make_euler_scene(n_sample, n_sample, oracle_kind=kind)
        was = o.set_ftz_daz(ftz)
        secs = []
        try:
            for _ in range(2):
                t0 = time.perf_counter()
                o.step(iters, 1 / 60.0, 1)
                secs.append(time.perf_counter() - t0)
        finally:
            if was >= 0:
                o.set_ftz_daz(was)
        interior = (o.num_x - 2) * (o.num_y - 2)
        out[label] = {"step1_s": secs[0], "step2_s": secs[1], "step1_cell_updates_per_s": interior * iters / secs[0],
                      "step2_cell_updates_per_s": interior * iters / secs[1]}
        o.close()
    out["value"] = out["ieee"]["step2_cell_updates_per_s"]
    return out


def sharded_solver_bench(rank, world, local_rank, nf=8192, rows_per_rank=1024, iters=200, reps=3):
    """Weak-scaling leg of BASELINE configs[4] (8192^2 with obstacle, 200 SOR sweeps: solver- and halo-bound): the grid is
    cut into `world` slabs of rows_per_rank rows; the velocity halos (9 rows) are exchanged with both neighbours every 8 colour passes over
    NCCL (9-row deep halo: one exchange per 8 colour passes).  8 ranks x 1024 rows == the 8192^2 grid of the config."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import simsandgames_b200 as sg
    from simsandgames_b200 import parallel
    ns = rows_per_rank * world
    uid = parallel.broadcast_unique_id() if world > 1 else None
    sh = sg.ShardedProjection(nf, ns, True, rank, world, unique_id=uid, device=local_rank)
    j0, j1 = sh.w0, sh.w1
    jj = np.arange(j0, j1, dtype=np.float32)[:, None]
    ii = np.arange(nf, dtype=np.float32)[None, :]
    s = np.ones((j1 - j0, nf), np.float32)
    s[:, 0] = 0; s[:, -1] = 0
    if j0 == 0: s[0] = 0
    if j1 == ns: s[-1] = 0
    disc = (ii + .5 - .75 * nf) ** 2 + (jj + .5 - .67 * ns) ** 2 < (.05 * nf) ** 2      # static obstacle (SURVEY S4 rule)
    s[disc] = 0
    rng = np.random.default_rng(1234 + rank)
    sh.upload_rows("s", s, j0)
    sh.upload_rows("cell_type", np.where(s == 0, 2, 0).astype(np.int32), j0)
    sh.upload_rows("vel_fast", rng.normal(0, 1, s.shape).astype(np.float32) * s, j0)
    sh.upload_rows("vel_slow", rng.normal(0, 1, s.shape).astype(np.float32) * s, j0)
    stream = torch.cuda.ExternalStream(sh.stream, device=torch.device("cuda", local_rank))
    cp = 1000.0 * (3.0 / nf) / (1 / 60.0 * 100 / nf)
    sh.solve(iters, cp, 1.9)   # warm-up
    sh.synchronize(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        sh.solve(iters, cp, 1.9)
    e1.record(stream)
    sh.synchronize(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    launches, exchanges = sh.counters()
    updates = float(nf - 2) * float(ns - 2) * iters
    out = {"workload": f"red-black SOR, {nf} x {ns} cells ({rows_per_rank} rows per GPU), {iters} sweeps, static disc obstacle (BASELINE configs[4] geometry at 8 GPUs)",
           "n_gpus": world, "ms_per_solve": ms, "cell_updates_per_s": updates / (ms * 1e-3),
           "alg_GBps_per_gpu": 36.0 * updates / world / (ms * 1e-3) / 1e9, "halo_exchanges_per_solve": exchanges // (reps + 1) if world > 1 else 0,
           "halo_rows": sh.halo, "halo_bytes_per_exchange_per_neighbour": 2 * sh.halo * 4 * nf, "scaling": "weak"}
    sh.close()
    return out


def eulerian_bench(local_rank, x=2048, y=2048, iters=100, steps=10):
    """BASELINE configs[1]: eulerian_fluid smoke 2D 2048^2 grid advect+project, 100 pressure iters, 1 B200."""
    import torch
    import simsandgames_b200 as sg
    sim = sg.scenes.make_wind_tunnel(x, y, device=local_rank)
    stream = torch.cuda.ExternalStream(sim.stream, device=torch.device("cuda", local_rank))
    dt = 1 / 60.0
    out = {}
    for name, order in (("lexicographic", sg.LEX_WAVEFRONT), ("red_black", sg.RED_BLACK)):
        sim.step(iters, dt, n_steps=3, solver_order=order)
        sim.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sim.step(iters, dt, n_steps=steps, solver_order=order)
        e1.record(stream)
        sim.synchronize(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        C = sim.getNumCells()
        interior = (sim.getNumX() - 2) * (sim.getNumY() - 2)
        out[name] = {"ms_per_step": ms, "cell_updates_per_s": interior * iters / (ms * 1e-3),
                     "alg_GBps": (86.0 + 25.0 * iters) * C / (ms * 1e-3) / 1e9}
    out["workload"] = f"eulerian_fluid wind tunnel {sim.getNumX()}x{sim.getNumY()} cells, {iters} pressure iters + extrapolate + advectVel + advectSmoke per step (BASELINE configs[1])"
    out["alg_bytes_formula"] = "(86 + 25*iters) * C  (SURVEY 8d)"
    sim.close()
    return out


def sharded_flip_bench(args, rank, world, local_rank):
    """N>1 headline: ONE slab-sharded FlipFluid over `world` GPUs (SURVEY 8d weak-scaling series): `world` dam-break tanks of
    BASELINE configs[2] stacked along y (the partition axis) and separated by solid rows, each rank owning one tank's rows.
    Every step runs the real exchanges: ghost particles, halo rows after P2G and after the solve, deep-halo red-black sweeps."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import simsandgames_b200 as sg
    from simsandgames_b200 import parallel, scenes
    f32 = np.float32
    N = args.grid
    sp = f32(3) / f32(N)
    width = float((f32(N) - f32(.5)) * sp)
    height = float((f32(N * world) - f32(.5)) * sp)
    radius = float(f32(.19) * sp)
    dt = float((f32(1) / f32(60)) * (f32(100) / f32(N)))
    tank_h = float((f32(N) - f32(.5)) * sp)
    pos, _, P0 = scenes.tank_particles(width, tank_h, float(sp), radius, want_s=False)
    uid = parallel.broadcast_unique_id()
    sim = sg.FlipFluidSharded(1000.0, width, height, float(sp), radius, int(P0 * 1.3), P0 * world, rank, world, unique_id=uid, device=local_rank)
    assert (sim.f_num_x, sim.f_num_y) == (N, N * world) and (sim.j0, sim.j1) == (rank * N, (rank + 1) * N)
    pos = pos.reshape(-1, 2)
    pos[:, 1] += f32(rank * N) * f32(sim.h)                      # this rank's tank sits in its own rows
    h_pos = torch.from_numpy(pos.reshape(-1)).pin_memory()
    h_ids = torch.arange(rank * P0, (rank + 1) * P0, dtype=torch.int32).pin_memory()
    w0, w1 = max(sim.j0 - sim.halo, 0), min(sim.j1 + sim.halo, sim.f_num_y)
    jj = np.arange(w0, w1)[:, None] % N
    s = np.ones((w1 - w0, N), np.float32)
    s[:, 0] = 0; s[:, -1] = 0
    s[(jj == 0)[:, 0] | (jj == N - 1)[:, 0]] = 0                 # floor and ceiling of every tank
    h_s = torch.from_numpy(s.reshape(-1)).pin_memory()
    h_zero = torch.zeros((w1 - w0) * N, dtype=torch.float32).pin_memory()
    kw = dict(dt=dt, gravity=9.81, flip_ratio=0.9, num_pressure_iters=50, num_particle_iters=2, over_relaxation=1.9, compensate_drift=True,
              separate_particles=True, obstacle_x=0.0, obstacle_y=0.0, obstacle_vel_x=0.0, obstacle_vel_y=0.0, obstacle_radius=0.0)
    stream = torch.cuda.ExternalStream(sim.stream, device=torch.device("cuda", local_rank))

    def barrier():
        sim.synchronize(); torch.cuda.synchronize(); dist.barrier()

    def reseed():
        for name in ("u", "v", "prev_u", "prev_v", "p", "particle_density"):
            sim.upload_rows_from(name, h_zero.data_ptr(), w0, w1 - w0)
        sim.upload_rows_from("s", h_s.data_ptr(), w0, w1 - w0)
        sim.particle_rest_density = 0.0
        sim.upload_particles_from(h_pos.data_ptr(), h_ids.data_ptr(), P0)
        sim.simulate(**kw, n_steps=args.warmup)

    SEG = 8   # red-black at 50 sweeps leaves the regime the reference stays in a little earlier than the lexicographic order
    ms, done = 0.0, 0
    launches = 0
    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.start()
    while done < args.steps:
        n = min(SEG, args.steps - done)
        reseed(); barrier()
        l0 = sim.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sim.simulate(**kw, n_steps=n)
        e1.record(stream)
        barrier()
        ms += e0.elapsed_time(e1); launches += sim.launch_count - l0; done += n
    clk = clocks.stop() if clocks else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    # end to end with host buffers: per step the caller's grid write (s rows) goes H2D, the owned particle positions come back D2H
    e2e_steps = max(3, min(args.steps, 6))
    h_out = torch.empty(2 * int(P0 * 1.3), dtype=torch.float32).pin_memory()
    reseed(); barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sim.upload_rows_from("s", h_s.data_ptr(), w0, w1 - w0)
        sim.simulate(**kw, n_steps=1)
        n_own = sim.readback_owned_pos_into(h_out.data_ptr(), int(P0 * 1.3))
    sim.synchronize()
    wall = time.perf_counter() - t0
    t = torch.tensor([wall * 1e3], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    cnt = torch.tensor([n_own], dtype=torch.int64, device="cuda"); dist.all_reduce(cnt)
    assert int(cnt.item()) == P0 * world, "particles lost or duplicated across ranks"
    phases = None
    if os.environ.get("FLIPGPU_BENCH_PHASES") == "1":
        # opt-in diagnostic pass (untimed): rank 0's CUDA-event phases of the sharded step; the exchanges are booked on
        # the phase they feed (particles -> integrate_bin, post-P2G rows -> p2g, post-solve rows -> g2p)
        reseed(); barrier()
        sim.enable_timings(True); sim.reset_timings()
        sim.simulate(**kw, n_steps=SEG)
        barrier()
        phases = {k: v / SEG for k, v in sim.timings().items()}
        sim.enable_timings(False)
    j0, j1, halo, band, exchanges = sim.shard_info()
    info = dict(phase_ms_per_step=phases, ms_total=ms_max, P_global=P0 * world, P_per_gpu=P0, launches=launches, e2e_ms=e2e_ms, e2e_steps=e2e_steps,
                h2d=int(h_s.numel() * 4), d2h=int(n_own * 8), halo_rows=halo, ghost_band_rows=band, nccl_exchanges=exchanges,
                grid=[N, N * world], clocks=clk)
    sim.close()
    return info


def shard_selfcheck(rank, world, local_rank, steps=6):
    """Before anything is timed: the default tank (BASELINE configs[0] scene) stepped by the slab-sharded path over `world` ranks must
    equal the one-GPU red-black run bit for bit -- every particle, u, v, p and the cell types.  Returns "bit-identical" or raises."""
    import numpy as np
    import torch.distributed as dist
    import simsandgames_b200 as sg
    from simsandgames_b200 import parallel
    one, kw, g = sg.scenes.make_flip_tank("default", device=local_rank)
    P = one.num_particles
    pos0 = one.readback("particle_pos").reshape(-1, 2).copy()
    s0 = one.readback("s").copy()
    rest = None
    if rank == 0:
        one.simulate(**kw, n_steps=1, solver_order=sg.RED_BLACK, update_colors=False)   # latch the rest density on one GPU, then restart
        rest = float(one.particle_rest_density)
        one.upload("particle_pos", pos0.reshape(-1)); one.upload("particle_vel", np.zeros(2 * P, np.float32))
        for n in ("u", "v", "prev_u", "prev_v", "p", "particle_density"):
            one.upload(n, np.zeros(one.f_num_cells, np.float32))
        one.simulate(**kw, n_steps=steps, solver_order=sg.RED_BLACK, update_colors=False)
    box = [rest]
    dist.broadcast_object_list(box, src=0)
    rest = box[0]
    uid = parallel.broadcast_unique_id()
    sh = sg.FlipFluidSharded(g["density"], g["width"], g["height"], g["spacing"], g["radius"], 2 * P, P, rank, world, unique_id=uid, device=local_rank)
    sh.upload("s", s0)
    sh.particle_rest_density = rest
    mine = sh.owns(pos0)
    sh.upload_particles(pos0[mine], None, np.flatnonzero(mine).astype(np.int32))
    sh.simulate(**kw, n_steps=steps)
    p_, v_, i_ = sh.readback_particles()
    part = dict(pos=p_, vel=v_, ids=i_, u=sh.readback_own_rows("u"), v=sh.readback_own_rows("v"), p=sh.readback_own_rows("p"), ct=sh.readback_own_rows("cell_type"))
    parts = [None] * world
    dist.all_gather_object(parts, part)
    sh.close()
    verdict = "bit-identical"
    if rank == 0:
        ids = np.concatenate([q["ids"] for q in parts])
        ok = np.array_equal(np.sort(ids), np.arange(P))
        if ok:
            pos = np.empty((P, 2), np.float32); vel = np.empty((P, 2), np.float32)
            pos[ids] = np.concatenate([q["pos"] for q in parts]).reshape(-1, 2)
            vel[ids] = np.concatenate([q["vel"] for q in parts]).reshape(-1, 2)
            ok = (np.array_equal(pos.view(np.int32).reshape(-1), one.readback("particle_pos").view(np.int32)) and
                  np.array_equal(vel.view(np.int32).reshape(-1), one.readback("particle_vel").view(np.int32)))
            for key, name in (("u", "u"), ("v", "v"), ("p", "p"), ("ct", "cell_type")):
                got = np.concatenate([q[key] for q in parts], axis=0).reshape(-1)
                ok = ok and np.array_equal(got.view(np.int32), one.readback(name).view(np.int32))
        verdict = "bit-identical" if ok else "MISMATCH"
    one.close()
    box = [verdict]
    dist.broadcast_object_list(box, src=0)
    if box[0] != "bit-identical":
        raise SystemExit(f"bench.py: the {world}-rank sharded step differs from the one-GPU red-black run on the default tank -- no line is printed")
    return box[0]


def synthetic_rank_lattice(N, j0, j1, build=True, row_histogram=False):
    """The particles of the synthetic dam-break (SURVEY 8d rule; lattice of flip_fluid/src/main.cpp:47-66) whose cell row lies
    in grid rows [j0, j1), with their global caller indices p = i*num_y + j -- without materialising the other ranks' particles.
    Same fp32 expressions as flip_scene_tank and the device ownership test."""
    import numpy as np
    f32 = np.float32
    sp = f32(3) / f32(N)
    W = (f32(N) - f32(.5)) * sp
    r = f32(.19) * sp
    dt = float((f32(1) / f32(60)) * (f32(100) / f32(N)))
    dx = f32(2) * r
    dy = f32(np.sqrt(3.0) / 2 * float(dx))
    num_x = int((f32(.6) * W - f32(2) * sp - f32(2) * r) / dx)
    num_y = int((f32(.8) * W - f32(2) * sp - f32(2) * r) / dy)
    jj = np.arange(num_y)
    y = (sp + r) + dy * jj.astype(np.float32)
    h_grid = max(W / f32(N), W / f32(N))                # flip_fluid.h:73 with f_num_x = f_num_y = N
    row = np.clip(((f32(1) / h_grid) * y).astype(np.int32), 0, N - 1)
    mine = jj[(row >= j0) & (row < j1)]
    out = dict(sp=sp, W=W, r=r, dt=dt, num_x=num_x, num_y=num_y, P_global=num_x * num_y, n_own=int(mine.size) * num_x, h_grid=h_grid)
    if row_histogram:
        out["row_particles"] = np.bincount(row, minlength=N).astype(np.float64) * num_x
    if build:
        ii = np.arange(num_x, dtype=np.float32)[:, None]
        pos = np.empty((num_x, mine.size, 2), np.float32)
        pos[..., 0] = ((sp + r) + dx * ii) + np.where(mine % 2 == 0, f32(0), r).astype(np.float32)[None, :]
        pos[..., 1] = y[mine][None, :]
        out["pos"] = pos.reshape(-1, 2)
        out["ids"] = (np.arange(num_x, dtype=np.int64)[:, None] * num_y + mine[None, :]).astype(np.int32).reshape(-1)
    return out


def single_tank_sharded(args, rank, world, local_rank, N, with_obstacle=False, steps=None):
    """ONE flip_fluid dam-break of N^2 cells (SURVEY 8d rule), slab-sharded over `world` GPUs -- a fixed global problem, so ranks
    holding air own few particles.  BASELINE configs[3] (N=16384, 1 029 824 840 particles, 50 sweeps) and configs[4] (N=8192 with
    the disc obstacle, 200 sweeps).  Every rank builds only its own lattice rows.  Returns the result dict on rank 0 (None elsewhere)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import simsandgames_b200 as sg
    from simsandgames_b200 import parallel
    steps = steps or args.steps
    iters = 200 if with_obstacle else 50
    uid = parallel.broadcast_unique_id()
    # slabs balanced by estimated work per grid row: particles x (cost per particle) + cells x (cost per cell and sweep), both
    # taken from the one-GPU phase timings at S2 (8.4 ms per 64.3 M particles, 4.4 ms per 16.8 M cells x 50 sweeps)
    rows_p = synthetic_rank_lattice(N, 0, N, build=False, row_histogram=True)["row_particles"]
    weights = rows_p * 1.31e-7 + N * (iters * 5.2e-9 + 2.0e-8)
    row_starts = parallel.weighted_partition(weights, world, min_rows=16)
    j0, j1 = row_starts[rank], row_starts[rank + 1]
    geo = synthetic_rank_lattice(N, j0, j1, build=False)
    sp, W, r, dt, num_x, num_y, P_global, n_own, h_grid = (geo[k] for k in ("sp", "W", "r", "dt", "num_x", "num_y", "P_global", "n_own", "h_grid"))
    cap = int(max(n_own * 1.25, 4.0e6)) + 2 * num_x * 64
    sim = sg.FlipFluidSharded(1000.0, float(W), float(W), float(sp), float(r), cap, P_global, rank, world, unique_id=uid, device=local_rank,
                              row_starts=row_starts)
    assert (sim.f_num_x, sim.f_num_y) == (N, N) and (sim.j0, sim.j1) == (j0, j1) and np.float32(sim.h) == h_grid, (sim.f_num_x, sim.j0, sim.h, h_grid)
    built = synthetic_rank_lattice(N, j0, j1, build=True)
    pos, ids = built["pos"], built["ids"]
    assert sim.owns(pos[:1000]).all()
    sim.upload_particles(pos.reshape(-1), None, ids)
    del pos, ids
    w0, w1 = max(j0 - sim.halo, 0), min(j1 + sim.halo, N)
    s = np.ones((w1 - w0, N), np.float32)
    s[:, 0] = 0; s[:, -1] = 0
    if w0 == 0: s[0] = 0
    if w1 == N: s[-1] = 0
    sim.upload_rows("s", s, w0)
    del s
    ox = oy = orad = 0.0
    if with_obstacle:   # static disc of the S4 rule through the device setObstacle (main.cpp:104-122 incl. its loop bounds)
        f32 = np.float32
        ox, oy, orad = float(f32(.75) * W), float(f32(.67) * W), float(f32(.05) * W)
        sim.set_obstacle(ox, oy, 0.0, 0.0, orad)
    kw = dict(dt=dt, gravity=9.81, flip_ratio=0.9, num_pressure_iters=iters, num_particle_iters=2, over_relaxation=1.9, compensate_drift=True,
              separate_particles=True, obstacle_x=ox, obstacle_y=oy, obstacle_vel_x=0.0, obstacle_vel_y=0.0, obstacle_radius=orad)
    stream = torch.cuda.ExternalStream(sim.stream, device=torch.device("cuda", local_rank))
    sim.simulate(**kw, n_steps=args.warmup)
    sim.synchronize(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sim.simulate(**kw, n_steps=steps)
    e1.record(stream)
    sim.synchronize(); torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    phases = None
    if os.environ.get("FLIPGPU_BENCH_PHASES") == "1":   # extra untimed pass: every rank's CUDA-event phases (exchanges booked on the phase they feed)
        sim.enable_timings(True); sim.reset_timings()
        sim.simulate(**kw, n_steps=2)
        sim.synchronize(); torch.cuda.synchronize(); dist.barrier()
        mine = {k: round(v / 2, 3) for k, v in sim.timings().items()}
        sim.enable_timings(False)
        phases = [None] * world
        dist.all_gather_object(phases, mine)
    h_out = torch.empty(2 * cap, dtype=torch.float32).pin_memory()
    n_now = sim.readback_owned_pos_into(h_out.data_ptr(), cap)
    cnt = torch.tensor([n_now, n_own], dtype=torch.int64, device="cuda"); dist.all_reduce(cnt)
    owned = [None] * world
    dist.all_gather_object(owned, n_now)
    assert int(cnt[0].item()) == P_global == int(cnt[1].item()), "particles lost or duplicated across ranks"
    mem = torch.cuda.mem_get_info(local_rank)
    sim.close()
    del h_out
    if rank != 0:
        return None
    peak, peak_src = peaks()
    s_per_step = ms * 1e-3 / steps
    step_bytes = 140.0 * P_global + (96.0 + 36.0 * iters) * float(N) * N
    name = "configs[4]" if with_obstacle else "configs[3]"
    return {"metric": METRIC, "value": P_global / s_per_step, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "phase_ms_per_step_by_rank": phases,
            "cell_updates_per_s_whole_step": float(N - 2) * (N - 2) * iters / s_per_step,
            "config": {"workload": f"flip_fluid synthetic dam-break {N}^2 grid, {P_global} particles, {iters} SOR iters" +
                                   (", static disc obstacle" if with_obstacle else "") + f", slab-sharded (BASELINE {name})",
                       "grid": [N, N], "particles_per_gpu": owned, "slab_row_starts": [int(x) for x in row_starts], "solver_order": "red_black",
                       "parallelism": f"y-slabs over {world} GPUs (balanced by estimated work per row), NCCL halo/ghost exchange",
                       "gpu_mem_used_GB_rank0": round((mem[1] - mem[0]) / 1e9, 1)},
            "roofline": {"bound": "hbm", "kernel": "whole sharded step (all GPUs)", "achieved": step_bytes / s_per_step / 1e9 / world, "peak": peak,
                         "unit": "GB/s per GPU", "frac": step_bytes / s_per_step / 1e9 / world / peak, "traffic": None, "peak_source": peak_src}}


def s3_bench(args, rank, world, local_rank, N=16384):
    import torch.distributed as dist
    line = single_tank_sharded(args, rank, world, local_rank, N)
    if rank == 0:
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def s0_default_tank(local_rank, steps=1000):
    """BASELINE configs[0]: the reference's own default tank (134 x 101 cells, 19 092 particles), 1000 steps -- the launch-bound end of
    the range: the whole step replayed as one CUDA graph vs queued launch by launch (the CPU reference takes 17.5 ms per step here)."""
    import torch
    import simsandgames_b200 as sg
    out = {"workload": "flip_fluid default tank 134x101 cells, 19092 particles, simulate(1/60, 9.81, .9, 50, 2, 1.9, true, true, ...), 1000 steps (BASELINE configs[0])"}
    for label, graphs in (("cuda_graph", True), ("launch_by_launch", False)):
        sim, kw, geo = sg.scenes.make_flip_tank("default", device=local_rank)
        sim.set_graph_mode(graphs)
        stream = torch.cuda.ExternalStream(sim.stream, device=torch.device("cuda", local_rank))
        sim.simulate(**kw, n_steps=20, update_colors=False)
        sim.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        sim.simulate(**kw, n_steps=steps, update_colors=False)
        e1.record(stream)
        sim.synchronize(); torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 1e3 * wall) / steps
        out[label] = {"ms_per_step": ms, "steps_per_s": 1e3 / ms, "particle_steps_per_s": sim.num_particles * 1e3 / ms, "graph_replays": sim.graph_replays}
        sim.close()
    return out


def developed_flow(local_rank, N=BENCH_N, steps=6):
    """The same S2 scene in a DEVELOPED state: seeded velocity noise (sigma 0.6, ~1.7 radii of random motion per step) on top of the
    lattice, three steps to let it splash -- crowded hash cells (up to ~10 particles), multi-cell moves, no lattice regularity -- then
    `steps` timed steps.  The headline times steps 4-15 of the dam-break, a near-lattice state; this leg says what disorder costs."""
    import numpy as np
    import torch
    import simsandgames_b200 as sg
    sim, kw, geo = sg.scenes.make_flip_tank("synthetic", N=N, device=local_rank)
    P = sim.num_particles
    rng = np.random.default_rng(20261020)
    sim.upload("particle_vel", rng.normal(0.0, 0.6, 2 * P).astype(np.float32))
    stream = torch.cuda.ExternalStream(sim.stream, device=torch.device("cuda", local_rank))
    sim.simulate(**kw, n_steps=3, update_colors=False)
    sim.synchronize()
    sim.enable_timings(True); sim.reset_timings()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sim.simulate(**kw, n_steps=steps, update_colors=False)
    e1.record(stream)
    sim.synchronize(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ph = {k: v / steps for k, v in sim.timings().items()}
    sim.enable_timings(False)
    cnt = sim.readback("num_cell_particles")
    out = {"workload": f"S2 scene with seeded velocity noise (sigma 0.6), steps 4-{3 + steps}", "ms_per_step": ms, "value": P / (ms * 1e-3), "unit": UNIT,
           "phase_ms_per_step": ph, "max_particles_per_hash_cell": int(cnt.max()), "hash_cells_with_3_or_more": int((cnt >= 3).sum())}
    sim.close()
    return out


def s4_single_gpu(local_rank, steps=4, warmup=3, N=8192):
    """BASELINE configs[4] on ONE GPU: 8192^2 cells with the disc obstacle, 257 311 935 particles, 200 SOR iterations per step."""
    import torch
    import simsandgames_b200 as sg
    sim, kw, geo = sg.scenes.make_flip_tank("synthetic", N=N, with_obstacle=True, device=local_rank)
    P, nx, ny, iters = sim.num_particles, sim.f_num_x, sim.f_num_y, kw["num_pressure_iters"]
    stream = torch.cuda.ExternalStream(sim.stream, device=torch.device("cuda", local_rank))
    sim.simulate(**kw, n_steps=warmup, update_colors=False)
    sim.synchronize()
    sim.enable_timings(True); sim.reset_timings()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sim.simulate(**kw, n_steps=steps, update_colors=False)
    e1.record(stream)
    sim.synchronize(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ph = {k: v / steps for k, v in sim.timings().items()}
    sim.enable_timings(False)
    assert bool((sim.readback("particle_pos")[:4096] == sim.readback("particle_pos")[:4096]).all())   # finite
    sim.close()
    peak, _ = peaks()
    step_bytes = 140.0 * P + (96.0 + 36.0 * iters) * float(nx) * ny
    return {"workload": f"flip_fluid synthetic dam-break {N}^2 grid with static disc obstacle, {P} particles, {iters} SOR iters (BASELINE configs[4]), 1 GPU, steps {warmup + 1}-{warmup + steps}",
            "ms_per_step": ms, "value": P / (ms * 1e-3), "unit": UNIT, "cell_updates_per_s": float(nx - 2) * (ny - 2) * iters / (ph["solve"] * 1e-3),
            "phase_ms_per_step": ph, "solver_order": "lexicographic (blocked wavefront)",
            "step_roofline": {"alg_bytes_per_step": step_bytes, "achieved_GBps": step_bytes / (ms * 1e-3) / 1e9, "frac_of_measured": step_bytes / (ms * 1e-3) / 1e9 / peak}}


def reference_arm(args, rank, world):
    if rank != 0:
        return
    n_sample = cpu_sample_size(args.steps, args.warmup)
    r = run_cpu_reference(n_sample, args.steps, args.warmup)
    sample = (f"same dam-break rule at {n_sample}^2 cells / {r['particles']} particles, {args.steps} steps after {args.warmup} warm-up, "
              f"1 thread (the reference is serial: no threads/OpenMP/SIMD in first-party code)")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"flip_fluid synthetic dam-break {n_sample}^2 grid, {r['particles']} particles, 50 SOR iters, 2 separation sweeps: a bounded CPU "
                               f"sample of BASELINE configs[2] (same scene rule; the full 4096^2 / 64264080-particle step takes the serial reference "
                               f"over a minute per step -- bench.py's product arm times one such step as cpu_baseline.full_size)",
                   "same_config_as_product_arm": False, "grid": [n_sample, n_sample], "particles": r["particles"],
                   "timing": "host steady clock around the reference's simulate()"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": 1, "kind": r["kind"], "sample": sample,
                         "cell_updates_per_s": r["cell_updates_per_s"], "phase_share": r["phase_share"],
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def product_arm(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    import simsandgames_b200 as sg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if world > 1 and args.s3:
        s3_bench(args, rank, world, local_rank, N=16384 if args.grid == BENCH_N else args.grid)
        return
    if world > 1:
        selfcheck = shard_selfcheck(rank, world, local_rank)      # fails the run (no line) unless bit-identical
        info = sharded_flip_bench(args, rank, world, local_rank)
        shard = sharded_solver_bench(rank, world, local_rank)
        extra = {}
        if not args.no_extra_legs:
            # BASELINE configs[4] as a full sharded step (8192^2 + obstacle, 200 sweeps), and configs[3] (16384^2, 1.03e9 particles) where it fits
            for key, n_grid, obst, min_world in (("s4_8192_obstacle_sharded", 8192, True, 2), ("s3_16384_sharded", 16384, False, 4)):
                if world < min_world:
                    continue
                try:
                    extra[key] = single_tank_sharded(args, rank, world, local_rank, n_grid, with_obstacle=obst, steps=4)
                except SystemExit:
                    raise
                except Exception as exc:
                    extra[key] = {"error": str(exc)[:300]}
        if rank == 0:
            peak, peak_src = peaks()
            N = args.grid
            s_per_step = info["ms_total"] * 1e-3 / args.steps
            step_bytes = 140.0 * info["P_global"] + (96.0 + 36.0 * 50) * float(N) * N * world
            line = {
                "metric": METRIC, "value": info["P_global"] / s_per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": info["ms_total"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"{world} x ({WORKLOAD}) as ONE slab-sharded FlipFluid: the tanks are stacked along y and separated by solid rows (SURVEY 8d weak-scaling series)",
                           "grid": info["grid"], "particles_per_gpu": info["P_per_gpu"], "pressure_iters": 50, "separation_sweeps": 2,
                           "solver_order": "red_black (the sharded order; the 1-GPU headline uses the lexicographic wavefront, which does not shard)",
                           "parallelism": f"y-slabs over {world} GPUs, one process per GPU; per step: 1 ghost-particle exchange, halo rows after P2G and after the solve, "
                                          f"{info['halo_rows']}-row deep-halo red-black sweeps (1 exchange per {info['halo_rows'] - 1} colour passes), all NCCL send/recv over NVLink",
                           "halo_rows": info["halo_rows"], "ghost_band_rows_last_step": info["ghost_band_rows"], "nccl_exchanges_total": info["nccl_exchanges"],
                           "reseed": f"state re-seeded (untimed, +{args.warmup} warm-up steps) every 8 timed steps (profiles/r1_stability.md)",
                           "l2": "working set ~4.7 GB per GPU per step >> 126 MB L2", "colours": "off",
                           "timing": "CUDA events on the library stream, barrier+synchronize both sides, max over ranks"},
                "clocks": dict({k: info["clocks"].get(k) for k in ("sm_mhz", "sm_max_mhz", "reasons", "samples", "power_w_max")}, note="GPU 0, sampled over the seed + timed segments"),
                "e2e": {"value": info["P_global"] * info["e2e_steps"] / (info["e2e_ms"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": info["h2d"],
                        "d2h_bytes_per_step": info["d2h"], "steps": info["e2e_steps"],
                        "what": "per rank and step: flip_upload_rows(s) H2D + sharded flip_step + flip_readback_particles(owned positions) D2H, pinned host buffers, host wall clock, max over ranks"},
                "gpu_launches": int(info["launches"]),
                "roofline": {"bound": "hbm", "kernel": "whole sharded step (per GPU)", "achieved": step_bytes / world / s_per_step / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": step_bytes / world / s_per_step / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                             "formula": "140*P + (96+36*iters)*C per GPU (SURVEY 8d)"},
                "sharded_solver": shard,
                "shard_selfcheck": selfcheck,
                "shard_selfcheck_what": f"default tank, 6 steps: {world}-rank sharded run vs one-GPU red-black run, particle_pos/vel, u, v, p, cell_type compared bitwise before the timed region",
            }
            line.update({k: v for k, v in extra.items() if v is not None})
            if info.get("phase_ms_per_step"):
                line["phase_ms_per_step"] = info["phase_ms_per_step"]
            print(json.dumps(line), flush=True)
        dist.destroy_process_group()
        return
    N = args.grid
    sim, kw, geo = sg.scenes.make_flip_tank("synthetic", N=N, device=local_rank)
    P, nx, ny = sim.num_particles, sim.f_num_x, sim.f_num_y
    iters = kw["num_pressure_iters"]
    stream = torch.cuda.ExternalStream(sim.stream, device=torch.device("cuda", local_rank))
    # initial state kept in pinned host memory: the scene is re-seeded (untimed) every SEG timed steps, because the
    # SURVEY dam-break rule (dt=(1/60)(100/N), 50 sweeps) diverges IN THE REFERENCE ITSELF after ~15 steps at N>=2048
    # (profiles/r1_stability.md); every timed step therefore runs on a state the reference also produces.
    SEG = 12
    h_pos0 = torch.from_numpy(sim.readback("particle_pos")).pin_memory()
    h_zero_p = torch.zeros(2 * P, dtype=torch.float32).pin_memory()
    h_zero_c = torch.zeros(nx * ny, dtype=torch.float32).pin_memory()
    h_s = torch.from_numpy(sim.readback("s")).pin_memory()

    def reseed():
        sim.num_particles = P
        sim.upload_from("particle_pos", h_pos0.data_ptr(), 2 * P)
        sim.upload_from("particle_vel", h_zero_p.data_ptr(), 2 * P)
        for name in ("u", "v", "prev_u", "prev_v", "p", "particle_density"):
            sim.upload_from(name, h_zero_c.data_ptr(), nx * ny)
        sim.upload_from("s", h_s.data_ptr(), nx * ny)
        sim.particle_rest_density = 0.0
        sim.simulate(**kw, n_steps=args.warmup, update_colors=False)   # untimed warm-up after every (re)seed

    def barrier():
        sim.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident throughput ("value"): K x flip_step, state resident in HBM
    clocks = ClockSampler(local_rank)
    clocks.start()                       # nvidia-smi needs ~0.3 s to deliver its first sample: start it before the warm-up
    sim.simulate(**kw, n_steps=args.warmup, update_colors=False)
    barrier()
    time.sleep(0.4)
    sim.simulate(**kw, n_steps=min(args.warmup, 3), update_colors=False)   # second warm-up burst: the clocks are ramped when timing starts
    barrier()
    reseed()
    barrier()
    sim.enable_timings(True)
    sim.reset_timings()
    launches0 = sim.launch_count
    ms, done, launches = 0.0, 0, 0
    while done < args.steps:
        n = min(SEG, args.steps - done)
        if done > 0:
            sim.enable_timings(False)
            reseed()
            barrier()
            sim.enable_timings(True)
            launches0 = sim.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sim.simulate(**kw, n_steps=n, update_colors=False)
        e1.record(stream)
        barrier()
        ms += e0.elapsed_time(e1)
        launches += sim.launch_count - launches0
        done += n
    clk = clocks.stop()
    phases = sim.timings()
    sim.enable_timings(False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    # ---- for information: the same step with the red-black order (north_star's named order; temporally blocked kernel).
    #      Not the headline: at the reference's 50 sweeps red-black leaves the regime the reference stays in earlier than
    #      the reference's own order (profiles/r1_stability.md), hence the shorter segment.  Guarded: never breaks the line.
    rb_variant = None
    try:
        if world == 1:
            rb_steps = 8
            reseed_rb = dict(kw)
            sim.num_particles = P
            sim.upload_from("particle_pos", h_pos0.data_ptr(), 2 * P)
            sim.upload_from("particle_vel", h_zero_p.data_ptr(), 2 * P)
            for name in ("u", "v", "prev_u", "prev_v", "p", "particle_density"):
                sim.upload_from(name, h_zero_c.data_ptr(), nx * ny)
            sim.upload_from("s", h_s.data_ptr(), nx * ny)
            sim.particle_rest_density = 0.0
            sim.simulate(**reseed_rb, n_steps=args.warmup, update_colors=False, solver_order=sg.RED_BLACK)
            barrier()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(stream)
            sim.simulate(**reseed_rb, n_steps=rb_steps, update_colors=False, solver_order=sg.RED_BLACK)
            r1.record(stream)
            barrier()
            rb_ms = r0.elapsed_time(r1) / rb_steps
            rb_variant = {"solver_order": "red_black (temporally blocked, 8 colour passes per launch)", "ms_per_step": rb_ms,
                          "value": P / (rb_ms * 1e-3), "unit": UNIT, "steps": rb_steps}
    except Exception as exc:   # diagnostic only
        rb_variant = {"error": str(exc)[:200]}

    # ---- end to end through the C ABI with HOST buffers (pinned): per step, the caller's per-frame writes go H2D
    #      (the s field that FluidUI::setObstacle rewrites every frame, main.cpp:106-122) and the array the caller
    #      reads every frame comes back D2H (particle_pos, main.cpp:173-176), in caller particle order.
    e2e_steps = max(8, args.steps)   # the last frame's PCIe copy is not hidden by a next step: it is part of the measurement
    h_frames = [torch.empty(2 * P, dtype=torch.float32).pin_memory() for _ in range(2)]   # double-buffered frames

    def frame_loop(n_frames):
        for k in range(n_frames):
            sim.upload_from_async("s", h_s.data_ptr(), h_s.numel())           # H2D: the caller's per-frame grid write (queued; h_s is never modified)
            sim.simulate(**kw, n_steps=1, update_colors=False)
            sim.readback_wait()                                                # frame k-1 has landed (its buffer is free again at k+1)
            sim.readback_pos_async(h_frames[k % 2].data_ptr(), 2 * P)          # D2H of frame k overlaps the simulation of frame k+1
        sim.readback_wait()

    reseed()
    frame_loop(3)                        # untimed warm-up of the frame pipeline (its staging buffers are allocated on first use)
    reseed()
    barrier()
    t0 = time.perf_counter()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(stream)
    frame_loop(e2e_steps)
    g1.record(stream)
    sim.synchronize()
    wall = time.perf_counter() - t0
    barrier()
    e2e_ms = max(g0.elapsed_time(g1), 1e3 * wall)  # device span vs host wall (which includes the last PCIe copy), whichever is longer
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms_max = float(t.item())
    h_pos = h_frames[(e2e_steps - 1) % 2]
    assert bool(torch.isfinite(h_pos[:1024]).all())

    # context for the e2e figure: what a bare pinned device->host copy of one frame achieves on this box (diagnostic only)
    d2h_link = None
    try:
        d_probe = torch.empty(2 * P, dtype=torch.float32, device=torch.device("cuda", local_rank))
        torch.cuda.synchronize()
        best = 0.0
        for _ in range(3):
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            h_frames[0].copy_(d_probe, non_blocking=True)
            c1.record()
            torch.cuda.synchronize()
            best = max(best, d_probe.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9)
        d2h_link = best
        del d_probe
    except Exception:   # never let a diagnostic break the bench line
        d2h_link = None

    sim.close()
    shard = sharded_solver_bench(rank, world, local_rank) if (world > 1 or args.with_shard_bench) else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    s_per_step = ms_max * 1e-3 / args.steps
    interior = (nx - 2) * (ny - 2)
    C = nx * ny
    # roofline of the dominant kernel = the blocked-wavefront SOR solve (one persistent launch per step)
    solve_ms_per_step = phases["solve"] / args.steps
    n_solver_launches = 1
    launch_s = solve_ms_per_step * 1e-3 / n_solver_launches
    alg_bytes_per_launch = 36.0 * interior * iters           # 36 B per cell-update x interior x iters per launch
    achieved = alg_bytes_per_launch / launch_s / 1e9
    step_bytes = 140.0 * P + (96.0 + 36.0 * iters) * C      # SURVEY 8d whole-step formula
    line = {
        "metric": METRIC, "value": world * P / s_per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD if N == BENCH_N else f"flip_fluid synthetic dam-break {N}^2 (non-default --grid)",
                   "grid": [nx, ny], "particles_per_gpu": P, "pressure_iters": iters, "separation_sweeps": 2,
                   "solver_order": "lexicographic (blocked wavefront, bit-identical to the reference order)", "colours": "off (render state, SURVEY 8d)",
                   "parallelism": "1 GPU",
                   "reseed": f"state re-seeded (untimed, +{args.warmup} warm-up steps) every {SEG} timed steps: the dam-break rule diverges in the reference itself after ~15 steps (profiles/r1_stability.md)",
                   "l2": "working set ~4.7 GB per step >> 126 MB L2 (no flush needed)",
                   "timing": "CUDA events on the library stream, barrier+synchronize both sides, max over ranks"},
        "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"],
                   "samples": clk["samples"], "power_w_max": clk.get("power_w_max")},
        "e2e": {"value": world * P * e2e_steps / (e2e_ms_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h_s.numel() * 4),
                "d2h_bytes_per_step": int(h_pos.numel() * 4), "steps": e2e_steps, "d2h_link_GBps_bare_copy": d2h_link,
                "what": "per frame: flip_upload_async(s) H2D (own stream, joined at the collision pass) + flip_step (its G2P kernel also writes the positions in caller order into a staging buffer) + flip_readback_pos_async(particle_pos) D2H into double-buffered pinned frames; the PCIe copy of frame k overlaps step k+1, every frame is waited for, the last frame's copy is not hidden"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "sor_gs_tiled_kernel (+ sor_prep_kernel; whole 50-iteration solve in one launch)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic_from_profile("sor_gs_tiled_kernel<(bool)1, (bool)1") if (args.grid == BENCH_N and world == 1) else None,
                     "traffic_source": f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel on this workload, read from {NCU_KERNELS_CSV} "
                                       "(written by scripts/ncu_summary.py from the committed capture, profiles/r2_ncu_flip_4096.md); temporal blocking keeps it ~150x below the "
                                       "algorithmic bytes, so frac > 1 is possible: step_roofline is the honest whole-step figure",
                     "peak_source": peak_src, "alg_bytes_per_launch": alg_bytes_per_launch, "launch_us": launch_s * 1e6, "launches_per_step": n_solver_launches,
                     "note": "achieved = the reference loop's cell visits x 36 B (SURVEY 8d) over the launch time: reference-equivalent work, not DRAM traffic. "
                             "The kernel skips all-air tiles and keeps a chunk of iterations in shared memory, and it is bound by the critical path of the "
                             "wavefront (latency), not by HBM -- see traffic, and step_roofline for the whole step"},
        "step_roofline": {"alg_bytes_per_step": step_bytes, "achieved": step_bytes / s_per_step / 1e9, "unit": "GB/s",
                          "frac_of_measured": step_bytes / s_per_step / 1e9 / peak, "frac_of_8TBs": step_bytes / s_per_step / 8e12,
                          "formula": "140*P + (96+36*iters)*C  (SURVEY 8d)"},
        "cell_updates_per_s": world * interior * iters / (solve_ms_per_step * 1e-3),
        "phase_ms_per_step": {k: v / args.steps for k, v in phases.items()},
    }
    if rb_variant is not None:
        line["red_black_variant"] = rb_variant
    if shard is not None:
        line["sharded_solver"] = shard
    if world == 1:
        line["eulerian_s1"] = eulerian_bench(local_rank)
    if world == 1 and not args.no_extra_legs and N == BENCH_N:
        try:
            line["developed_flow"] = developed_flow(local_rank)
        except Exception as exc:
            line["developed_flow"] = {"error": str(exc)[:300]}
        try:
            line["s0_default_tank"] = s0_default_tank(local_rank)
        except Exception as exc:
            line["s0_default_tank"] = {"error": str(exc)[:300]}
        try:
            line["s4_8192_obstacle"] = s4_single_gpu(local_rank)
        except Exception as exc:
            line["s4_8192_obstacle"] = {"error": str(exc)[:300]}
    if world == 1 and not args.no_cpu_baseline:
        line["eulerian_s1"]["cpu_baseline"] = run_cpu_reference_euler()
        n_sample = 1024
        r = run_cpu_reference(n_sample, 3, 0)
        line["cpu_baseline"] = {
            "value": r["value"], "unit": UNIT, "cores": 1, "kind": r["kind"],
            "sample": f"same dam-break rule at {n_sample}^2 cells / {r['particles']} particles, 3 steps, 1 thread of {os.cpu_count()} host CPUs "
                      f"({r['seconds']:.1f} s; the reference is serial)",
            "cell_updates_per_s": r["cell_updates_per_s"], "phase_share": r["phase_share"]}
        if not args.no_full_size_cpu and N == BENCH_N:
            try:
                line["cpu_baseline"]["full_size"] = run_cpu_reference_full_size()
            except Exception as exc:   # e.g. not enough host memory: report, never break the line
                line["cpu_baseline"]["full_size"] = {"error": str(exc)[:200]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="flipgpu", choices=["flipgpu", "reference"])
    ap.add_argument("--grid", type=int, default=BENCH_N, help="cells per side of the synthetic dam-break (default: the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-size-cpu", action="store_true", help="skip the one full-size (4096^2) step of the CPU reference (~80 s)")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the S4 (8192^2 + obstacle) / S3 legs")
    ap.add_argument("--s3", action="store_true", help="N>1 only: BASELINE configs[3], one 16384^2 / 1.03e9-particle tank sharded over the ranks")
    ap.add_argument("--with-shard-bench", action="store_true", help="also run the sharded-solver leg at N=1 (always run for N>1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
    else:
        if args.warmup < 3:
            args.warmup = 3
        product_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
