"""CPU: the reference arm of bench.py (the reference's own CPU code on a bounded sample) prints one well-formed JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "flip_particle_steps_per_s" and line["unit"] == "particle-steps/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["value"] > 1e5
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "4096^2" in line["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_eulerian_cpu_baseline_reports_both_denormal_modes():
    """the Eulerian CPU baseline is timed with and without FTZ/DAZ; the calling thread's mode is restored."""
    sys.path.insert(0, ROOT)
    import bench
    from oracle import cpu_oracle as co
    probe = co.EulerOracle(4, 4, 1000.0, 0.01)
    before = probe.set_ftz_daz(0)
    probe.set_ftz_daz(before)
    r = bench.run_cpu_reference_euler(n_sample=48, iters=5)
    assert r["cores"] == 1 and r["kind"] in ("reference", "port") and r["unit"] == "cell-updates/s"
    for mode in ("ieee", "ftz_daz"):
        assert r[mode]["step1_s"] > 0 and r[mode]["step2_cell_updates_per_s"] > 0
    assert r["value"] == r["ieee"]["step2_cell_updates_per_s"]
    after = probe.set_ftz_daz(before)
    assert after == before
