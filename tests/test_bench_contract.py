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


def test_roofline_record_arithmetic_and_traffic_file():
    """The roofline object of the bench line: achieved = SURVEY 8(d) bytes x spheres / kernel time, frac = achieved / peak,
    traffic = the committed ncu capture of the same workload (configs[1], ~1 M spheres) and nothing for any other size."""
    sys.path.insert(0, ROOT)
    import bench
    rec = {"contacts_per_sphere": 8.45, "value": 3.4e9,
           "kernel_ms_per_timestep": {"k_step_begin": 0.015, "k_force_integrate": 0.277}}
    n = 997920
    r = bench.roofline_record(rec, n)
    peak, _ = bench.peaks()
    assert r["kernel"] == "k_force_integrate" and r["unit"] == "GB/s" and r["bound"] == "hbm"
    want = (160.0 + 32.0 * 8.45) * n / 0.277e-3 / 1e9
    assert abs(r["achieved"] - want) < 1e-6 * want and abs(r["frac"] - want / peak) < 1e-9
    assert abs(r["whole_step_frac"] - (176.0 + 32.0 * 8.45) * 3.4e9 / 1e9 / peak) < 1e-9
    with open(os.path.join(ROOT, "profiles", "force_traffic.json")) as f:
        t = json.load(f)
    assert r["traffic"] == t["dram_bytes_read"] + t["dram_bytes_write"] and r["traffic_source"]
    assert os.path.exists(os.path.join(ROOT, t["capture"].split(" ")[0])), "the capture the traffic figure cites is committed"
    # measured traffic is well above the algorithmic bytes (duplicate history, candidate ids): the line must show it
    assert 1.5 < r["traffic"] / ((160.0 + 32.0 * 8.45) * n) < 2.5
    assert bench.roofline_record(rec, 4000000)["traffic"] is None
