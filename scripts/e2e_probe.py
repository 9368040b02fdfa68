"""Where does the end-to-end frame loop spend its time?  (diagnostic, GPU box)
Runs bench.py's e2e loop (upload s -> step -> wait frame k-1 -> queue readback of frame k) with parts switched off.
usage: python scripts/e2e_probe.py [N=4096] [steps=8]"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import simsandgames_b200 as sg
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sim, kw, g = sg.scenes.make_flip_tank("synthetic", N=N)
P = sim.num_particles
h_s = torch.from_numpy(sim.readback("s")).pin_memory()
frames = [torch.empty(2 * P, dtype=torch.float32).pin_memory() for _ in range(2)]
sim.simulate(**kw, n_steps=3, update_colors=False)
sim.synchronize()
VARIANTS = (("step only", False, False), ("upload+step", True, False), ("step+readback", False, True), ("full", True, True), ("full again", True, True))
if os.environ.get("PROBE_VARIANTS"):
    VARIANTS = tuple(v for v in VARIANTS if v[0] in os.environ["PROBE_VARIANTS"].split(","))
for label, up, rb in VARIANTS:
    sim.synchronize(); torch.cuda.synchronize()
    sim.enable_timings(True); sim.reset_timings()
    marks = []
    t0 = time.perf_counter()
    for k in range(K):
        a = time.perf_counter()
        if up:
            sim.upload_from_async("s", h_s.data_ptr(), h_s.numel())
        b = time.perf_counter()
        sim.simulate(**kw, n_steps=1, update_colors=False)
        c = time.perf_counter()
        if rb:
            sim.readback_wait()
        d = time.perf_counter()
        if rb:
            sim.readback_pos_async(frames[k % 2].data_ptr(), 2 * P)
        e = time.perf_counter()
        marks.append((b - a, c - b, d - c, e - d))
    if rb:
        sim.readback_wait()
    sim.synchronize()
    wall = time.perf_counter() - t0
    avg = [1e3 * sum(m[i] for m in marks[2:]) / max(len(marks) - 2, 1) for i in range(4)]
    ph = {k2: round(v / K, 3) for k2, v in sim.timings().items()}
    sim.enable_timings(False)
    print(json.dumps({"variant": label, "ms_per_step": round(1e3 * wall / K, 3), "phase_ms": ph,
                      "host_ms": {"upload_call": round(avg[0], 3), "step_call": round(avg[1], 3), "wait_call": round(avg[2], 3), "readback_call": round(avg[3], 3)}}), flush=True)
sim.close()
