"""bench.py prints exactly ONE JSON line with the keys the driver reads.  The reference arm runs on CPU;
the product arm needs a GPU (tiny workload, seconds)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def run_bench(*args):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must hold exactly one line, got %d" % len(lines)
    return json.loads(lines[0])


def test_reference_arm_json_line():
    d = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--side", "1200", "--grid", "16",
                  "--ref-sample", "60")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "PM grid vectors/sec" and d["unit"] == "vectors/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"] and d["gpu_launches"] == 0


@pytest.mark.gpu
def test_product_arm_json_line():
    d = run_bench("--steps", "3", "--warmup", "3", "--side", "1500", "--grid", "30", "--cpu-sample", "200", "--no-configs")
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline", "parity"} <= set(d)
    assert d["n_gpus"] == 1 and d["scaling"] == "weak" and d["dtype"] == "u8" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["gpu_launches"] >= d["steps"]
    assert d["e2e"]["h2d_bytes_per_step"] > 2 * 1500 * 1500 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert 0 < d["e2e"]["value"] <= d["value"] * 1.05
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in d["roofline"] and key in d["roofline_fma"]
    assert d["roofline"]["bound"] in ("hbm", "tensor") and d["roofline_fma"]["bound"] == "fma"
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    p = d["parity"]
    assert p["nan_pattern_equal"] and p["position_angle_equal"] == p["compared"] == p["r_bit_equal"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["value"] > 0
    if d["cpu_baseline"]["kind"] == "reference":            # the reference copy travelled: parity is reported against it too
        v = p["vs_reference"]
        assert v["nan_pattern_equal"] and v["unexplained"] == 0 and v["max_abs_dr"] <= 1e-4
        assert {"max_abs_dh", "n_dh_above_1e-4", "tie_explained", "position_angle_exact"} <= set(v)
    assert "e2e_pageable" in d and d["e2e_pageable"]["value"] > 0
    assert d["configs"] is None or {"cfg1", "cfg3", "cfg4"} <= set(d["configs"])
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
