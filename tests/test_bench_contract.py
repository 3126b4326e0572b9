"""bench.py contract pieces that do not need a GPU: the reference arm (the CPU port of the Multicore step on all host
threads) prints ONE JSON line with the agreed keys, and ranks other than 0 stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True,
                       timeout=600, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_json_line():
    out = run_bench("--impl", "reference", "--spheres", "20000", "--steps", "2", "--warmup", "1")
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sphere-steps/sec" and d["unit"] == "sphere-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None


def test_reference_arm_other_ranks_are_silent():
    out = run_bench("--impl", "reference", "--gpus", "2", "--spheres", "20000", "--steps", "1", "--warmup", "1",
                    env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.strip() == ""
