"""bench.py's output contract: one JSON line with the keys the driver reads.  The reference arm runs on the CPU (here
and on the GPU box); the product arm needs a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline")


def run_bench(*args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line expected, got %d" % len(lines)
    return json.loads(lines[0])


def check_common(line):
    for k in BASE_KEYS:
        assert k in line, k
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["higher_is_better"] is True
    assert line["warmup"] >= 3 and "workload" in line["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in line["e2e"], k
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in line["cpu_baseline"], k
    assert line["cpu_baseline"]["kind"] in ("reference", "port")


def test_reference_arm_line():
    """`bench.py --impl reference`: the reference's CPU implementation (oracle/_ref when built, else the port)."""
    line = run_bench("--impl", "reference", "--workload", "c1", "--steps", "2", "--warmup", "1")
    check_common(line)
    assert line["impl"] == "reference" and line["n_gpus"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"] == line["cpu_baseline"]["value"]


@pytest.mark.gpu
def test_product_arm_line():
    """The product arm on a small fleet (64 robots of C4's shape) with the CPU leg of the single-robot configuration
    left out of the picture: every key of the contract, a roofline object, launches counted, clocks sampled."""
    line = run_bench("--robots", "64", "--steps", "4", "--warmup", "3", "--no-extra")
    check_common(line)
    assert "impl" not in line or line["impl"] != "reference"
    assert line["gpu_launches"] == 3 * line["steps"], line["gpu_launches"]
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["e2e"]["value"] != line["value"]
    roof = line["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in roof, k
    assert roof["bound"] == "hbm" and 0 < roof["frac"] < 1.5 and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    assert "sm_mhz" in line["clocks"] and "reasons" in line["clocks"]
    assert line["config"]["workload"].startswith("c4")
