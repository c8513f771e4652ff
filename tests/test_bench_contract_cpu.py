"""bench.py's reference arm (`--impl reference`: the oracle port on the host cores, bounded sample) prints ONE JSON
line with the contract's keys; runs here on the small c1 workload in a few seconds.  The libvcof arm needs a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                        "--steps", "1", "--warmup", "0", "--gpus", "2"], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, RANK="0"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "denoising_steps_per_sec" and d["unit"] == "steps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 2
    assert d["value"] > 0 and abs(d["value"] - 1000.0 / d["ms_per_step"]) < 1e-9 * d["value"] + 1e-12
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "extrapolated" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("c1:") and d["config"]["parallelism"] == "sp2"


def test_reference_arm_other_ranks_exit_quietly():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                        "--gpus", "2"], capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK="1"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_l2_note_is_computed_per_workload():
    sys.path.insert(0, ROOT)
    import bench
    assert "no flush needed" in bench.workload_config("c2", 1)["l2"]
    assert "no flush needed" in bench.workload_config("c2", 8)["l2"]
    assert "not a valid bench configuration" in bench.workload_config("c1", 1)["l2"]
