"""bench.py's reference arm (`--impl reference`: the oracle on the host cores)
runs without a GPU; this checks that it prints ONE JSON line with the keys the
driver's contract names.  The GPU arm's line is checked on the B200
(tests/test_gpu_parity.py::test_bench_line_contract)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
             "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
             "e2e", "cpu_baseline")


def check_line(line, reference):
    for k in BASE_KEYS:
        assert k in line, k
    assert line["metric"] == "bounded fits solved/sec (batched)"
    assert line["unit"] == "fits/s" and line["higher_is_better"] is True
    assert line["dtype"] == "f64" and line["data"] == "synthetic"
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["vs_baseline"] is None            # BASELINE.md publishes no number
    cb = line["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in cb, k
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1
    e2e = line["e2e"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in e2e, k
    if reference:
        assert line["impl"] == "reference"
        assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
        assert e2e["value"] == line["value"] == cb["value"]
    else:
        assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0
        assert line["gpu_launches"] > 0
        rf = line["roofline"]
        for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
            assert k in rf, k
        assert rf["bound"] == "hbm" and 0 < rf["frac"] < 1
        assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
        ck = line["clocks"]
        assert "sm_mhz" in ck and "sm_max_mhz" in ck and "reasons" in ck


def run_bench(args, timeout):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args,
                         capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_line():
    line = run_bench(["--impl", "reference", "--steps", "1", "--warmup", "0"], 600)
    check_line(line, reference=True)
    tall = line["tall"]
    assert tall["impl"] == "reference" and tall["unit"] == "iterations/s"
    assert tall["config"]["extrapolated"] is True


@pytest.mark.gpu
def test_gpu_arm_line():
    line = run_bench(["--steps", "2", "--warmup", "3", "--no-tall"], 900)
    check_line(line, reference=False)
    p = line["cpu_baseline"]["parity_vs_gpu"]
    assert p["status_equal"] == 1.0 and p["x_rel_max"] < 1e-8 and p["obj_rel_max"] < 1e-8
