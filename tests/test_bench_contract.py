"""bench.py contract pieces that run without a GPU: the reference arm prints exactly one JSON line with
the required keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--workload", "C1"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("distill fwd+bwd") and d["higher_is_better"] is True
    # the arm times the UNMODIFIED reference (oracle/_ref, or /root/reference in the build container) at the full shard
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "B=8 of 8 samples" in d["cpu_baseline"]["sample"] and "unmodified reference" in d["cpu_baseline"]["sample"]
    assert d["steps"] == 1 and d["warmup"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] and "model" not in d["config"] and d["vs_baseline"] is None


def test_non_rank0_reference_arm_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""
