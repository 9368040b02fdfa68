"""Do a pinned H2D copy and a pinned D2H copy on two streams overlap on this box?  (diagnostic, GPU box)"""
import torch, time, json
d_big = torch.empty(128 * 2**20, dtype=torch.float32, device="cuda")     # 512 MiB
d_small = torch.empty(16 * 2**20, dtype=torch.float32, device="cuda")    # 64 MiB
h_big = torch.empty(128 * 2**20, dtype=torch.float32).pin_memory()
h_small = torch.empty(16 * 2**20, dtype=torch.float32).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(which):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if "d2h" in which:
        with torch.cuda.stream(s1):
            h_big.copy_(d_big, non_blocking=True)
    if "h2d" in which:
        with torch.cuda.stream(s2):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(s2); d_small.copy_(h_small, non_blocking=True); e1.record(s2)
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0), (e0.elapsed_time(e1) if "h2d" in which else None)
for w in (("d2h",), ("h2d",), ("d2h", "h2d"), ("d2h", "h2d")):
    run(w)
    wall, h2d_ms = run(w)
    print(json.dumps({"copies": w, "wall_ms": round(wall, 3), "h2d_64MiB_ms_on_its_stream": None if h2d_ms is None else round(h2d_ms, 3)}))
