"""Print the interesting keys of a bench.py JSON line.  usage: python scripts/show_line.py file.json [key ...]"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
keys = sys.argv[2:] or ["n_gpus", "ms_per_step", "value", "shard_selfcheck", "phase_ms_per_step", "e2e", "roofline", "step_roofline", "cell_updates_per_s"]
for k in keys:
    if k in d:
        v = d[k]
        if isinstance(v, dict):
            v = {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if not isinstance(b, str) or len(b) < 60}
        print(k, "=", v)
for k in ("sharded_solver", "s4_8192_obstacle", "s4_8192_obstacle_sharded", "s3_16384_sharded", "eulerian_s1", "red_black_variant", "cpu_baseline"):
    if k in d and isinstance(d[k], dict):
        v = d[k]
        print(k, "=", {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if isinstance(b, (int, float)) or a in ("error", "lexicographic", "red_black")})
