"""The bench.py JSON line contract, checked on the committed snapshot of the last GPU run
(profiles/r01_bench_n1.json) and on the reference arm, which runs on CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"]


def test_committed_bench_line_has_every_contract_key():
    line = json.loads(open(os.path.join(ROOT, "profiles", "r01_bench_n1.json")).read().strip().splitlines()[-1])
    for k in REQUIRED:
        assert k in line, k
    assert line["unit"] == "frames/s" and line["higher_is_better"] is True and line["scaling"] == "weak"
    assert "workload" in line["config"] and "l2" in line["config"]
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    e = line["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= line["value"] * 1.02
    r = line["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6 and r["traffic"]
    c = line["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["sample"]
    assert line["gpu_launches"] > 0


def test_reference_arm_runs_on_cpu_and_prints_the_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "frames/s"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port"


def test_dit_roofline_helper_on_the_snapshot():
    """bench.dit_roofline: algorithmic sampler FLOPs over the sampler's share of the step."""
    import bench
    line = json.loads(open(os.path.join(ROOT, "profiles", "r01_bench_n1.json")).read().strip().splitlines()[-1])
    r = bench.dit_roofline(line["ms_per_step"], line["stage_ms_eager"], 1392.3, 32)
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s"
    assert abs(r["algorithmic_tflop"] - (3.394 * 32 + 0.465)) < 1e-9
    assert 0.0 < r["frac"] < 1.0 and abs(r["frac"] - r["achieved"] / 1392.3) < 1e-12
    assert r["ms"] < line["ms_per_step"]
    assert bench.dit_roofline(1.0, {}, 1000.0)["achieved"] > 0          # degenerate inputs do not raise
