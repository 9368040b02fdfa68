"""Separation cost of the three S2 states bench.py knows (diagnostic, GPU box): near-lattice dam-break (lexicographic and red-black
trajectories) and the developed-flow state.  usage: [FLIPGPU_LIB=...] python scripts/dev_probe.py"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
d = bench.developed_flow(0, steps=4)
print(json.dumps({"lib": os.environ.get("FLIPGPU_LIB", "")[-16:], "developed_ms": round(d["ms_per_step"], 3), "developed_sep": round(d["phase_ms_per_step"]["separate"], 3)}), flush=True)
