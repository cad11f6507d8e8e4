"""The JSON line bench.py prints is a contract with the driver.  This checks the recorded lines of
the last GPU runs (profiles/r1d_bench_*.json, written by bench.py itself) for the keys and the
internal consistency the contract asks for - a cheap guard against editing bench.py into a shape
the driver cannot read.  (bench.py cannot run here: no GPU.)"""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(name + " not recorded")
    with open(path) as f:
        return json.load(f)


def test_own_arm_line():
    d = _load("r1d_bench_1gpu.json")
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "gpu_launches", "clocks", "roofline",
                "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["data"] == "synthetic" and d["dtype"] == "f32" and "workload" in d["config"]
    assert d["gpu_launches"] > 0
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # algorithmic bytes of the dominant kernel per launch: (8 + 4) B per sample + the window
    assert abs(r["bytes_per_launch"] - (12 * 1024 * r["spectra_per_launch"] + 4 * 1024)) < 1
    assert abs(r["achieved"] - r["bytes_per_launch"] / (r["ms_per_launch"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 8 * 384 * 1024 * 1024 and e["d2h_bytes_per_step"] > 0
    # whole-job value = samples per step / time per step
    samples = 384 * 1024 * 1024
    assert abs(d["value"] - d["n_gpus"] * samples / (d["ms_per_step"] * 1e-3) / 1e6) < 1e-6 * d["value"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_reference_arm_line():
    d = _load("r1d_bench_reference_arm.json")
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    own = _load("r1d_bench_1gpu.json")
    assert d["metric"] == own["metric"] and d["unit"] == own["unit"]
