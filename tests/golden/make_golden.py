"""Generates tests/golden/*.json from the UNMODIFIED reference (oracle/_ref, built from /root/reference by
oracle/Makefile).  Run in the build container:  python tests/golden/make_golden.py
The reference has no golden vectors of its own; these pin the C port and the scene builders to it."""
import hashlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cpu_oracle as co

HERE = os.path.dirname(os.path.abspath(__file__))
FLIP_ARRAYS = ["particle_pos", "particle_vel", "particle_color", "u", "v", "p", "prev_u", "particle_density", "cell_type",
               "cell_color", "num_cell_particles", "first_cell_particle", "cell_particle_ids"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def flip_record(o):
    P = o.num_particles
    rec = {n: sha(getattr(o, n)) for n in FLIP_ARRAYS}
    rec["rest_density_bits"] = int(np.float32(o.particle_rest_density).view(np.int32))
    rec["pos_head"] = [float(x) for x in o.particle_pos[:16]]
    rec["mean_y"] = float(o.particle_pos[1:2 * P:2].astype(np.float64).mean())
    return rec


def flip_case(kind, N=None, checkpoints=(1, 5, 20)):
    o, p, g = co.make_flip_scene(kind, N=N, oracle_kind="ref")
    out = {"dims": dict(f_num_x=o.f_num_x, f_num_y=o.f_num_y, p_num_x=o.p_num_x, p_num_y=o.p_num_y, num_particles=o.num_particles,
                        h_bits=int(np.float32(o.h).view(np.int32)), p_inv_bits=int(np.float32(o.p_inv_spacing).view(np.int32))),
           "scene": {"particle_pos": sha(o.particle_pos), "s": sha(o.s)}, "params": p, "geom": g, "steps": {}}
    done = 0
    for c in checkpoints:
        o.simulate(**p, n_steps=c - done); done = c
        out["steps"][str(c)] = flip_record(o)
    return out


def euler_case(x, y, iters, checkpoints=(1, 10, 40)):
    o = co.make_euler_scene(x, y, oracle_kind="ref")
    out = {"dims": dict(num_x=o.num_x, num_y=o.num_y), "scene": {n: sha(getattr(o, n)) for n in ("u", "m", "solid")}, "iters": iters, "steps": {}}
    done = 0
    for c in checkpoints:
        o.step(iters, 1 / 60.0, n_steps=c - done); done = c
        out["steps"][str(c)] = {n: sha(getattr(o, n)) for n in ("u", "v", "m", "pressure")}
        out["steps"][str(c)]["m_mean"] = float(o.m.astype(np.float64).mean())
    return out


def fluid3d_case(x, y, z, iters, checkpoints=(1, 4)):
    o = co.make_fluid3d_scene(x, y, z, oracle_kind="ref")
    out = {"dims": dict(num_x=o.num_x, num_y=o.num_y, num_z=o.num_z), "scene": {n: sha(getattr(o, n)) for n in ("u", "v", "w", "m", "solid")},
           "iters": iters, "steps": {}}
    done = 0
    for c in checkpoints:
        o.step(iters, 1 / 60.0, n_steps=c - done); done = c
        out["steps"][str(c)] = {n: sha(getattr(o, n)) for n in ("u", "v", "w", "m", "pressure")}
        out["steps"][str(c)]["m_mean"] = float(o.m.astype(np.float64).mean())
    return out


if __name__ == "__main__":
    assert co.available("flip", "ref"), "build oracle/_ref first (make -C oracle) -- needs /root/reference"
    gold = {"flip_default": flip_case("default"), "flip_synthetic_64": flip_case("synthetic", N=64, checkpoints=(1, 5, 12)),
            "euler_96x72": euler_case(96, 72, 40), "euler_33x20": euler_case(33, 20, 7, checkpoints=(1, 5)),
            "fluid3d_30x22x26": fluid3d_case(30, 22, 26, 20), "fluid3d_12x9x17": fluid3d_case(12, 9, 17, 5, checkpoints=(1, 3))}
    with open(os.path.join(HERE, "reference_golden.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("wrote", os.path.join(HERE, "reference_golden.json"))
