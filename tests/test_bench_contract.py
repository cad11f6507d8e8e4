"""The JSON line bench.py prints is a contract with the driver.  This checks the recorded lines of
the last GPU runs (profiles/r2_bench_*.json, written by bench.py itself) for the keys and the
internal consistency the contract asks for - a cheap guard against editing bench.py into a shape
the driver cannot read.  (bench.py cannot run here: no GPU.)"""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PASS_SAMPLES = 384 * 1024 * 1024          # one pass = 1 s of 100 Msps IQ at overlap 4


def _load(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(name + " not recorded")
    with open(path) as f:
        return json.load(f)


def test_own_arm_line():
    d = _load("r2_bench_1gpu.json")
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "gpu_launches", "clocks", "roofline",
                "cpu_baseline", "e2e", "burst", "unfolded", "one_stream", "configs"):
        assert key in d, key
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["data"] == "synthetic" and d["dtype"] == "f32" and "workload" in d["config"]
    assert "model" not in d["config"]
    assert d["gpu_launches"] > 0
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # algorithmic bytes of the dominant kernel per launch: (8 + 4) B per sample + the window
    assert abs(r["bytes_per_launch"] - (12 * 1024 * r["spectra_per_launch"] + 4 * 1024)) < 1
    assert abs(r["achieved"] - r["bytes_per_launch"] / (r["ms_per_launch"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    # measured DRAM traffic of the same kernel: within 5 % of the algorithmic bytes (no wasted re-reads)
    assert 0.9 < r["traffic"] / r["bytes_per_launch"] < 1.05
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    # e2e: the like-for-like arm - one sink frame of 8 calls of 1024 x 1024 cf32 in, 4.5 MiB out
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 8 * 8 * 1024 * 1024
    assert e["d2h_bytes_per_step"] == 4 * (1024 * 1024 + 128 * 1024 + 4 * 1024)
    assert "fosphor_cl_" in e["api"] and "PAGEABLE" in e["api"]
    assert abs(e["value"] - d["n_gpus"] * 8 * 1024 * 1024 / (e["ms_per_frame"] * 1e-3) / 1e6) < 1e-3 * e["value"]
    # whole-job value = samples per step / time per step, over a timed region of about a second
    passes = d["details"]["passes_per_step"]
    assert abs(d["value"] - d["n_gpus"] * passes * PASS_SAMPLES / (d["ms_per_step"] * 1e-3) / 1e6) < 1e-6 * d["value"]
    assert d["details"]["timed_region_s"] >= 1.0 and d["details"]["wf_rows"] == 1024
    assert abs(d["details"]["timed_region_s"] - d["steps"] * d["ms_per_step"] * 1e-3) < 1e-6
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons", "power_w_max"}
    # a run that saw a thermal or hardware slowdown would have to be re-measured; the power cap is kept and noted
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["burst"]["value"] >= d["value"] * 0.98
    names = [c["config"] for c in d["configs"]]
    assert any("cfg3" in n for n in names) and any("cfg4" in n for n in names)
    assert sum("sweep" in n for n in names) == 6 and sum("cfg5 stress" in n for n in names) == 2
    for c in d["configs"]:
        assert 0 < c["frac"] < 1 and c["Msamples_per_s"] > 0


def test_reference_arm_line():
    d = _load("r2_bench_reference_arm.json")
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    own = _load("r2_bench_1gpu.json")
    assert d["metric"] == own["metric"] and d["unit"] == own["unit"]
    assert d["config"] == own["config"]            # same_config: both arms print the very same dict
    assert d["higher_is_better"] == own["higher_is_better"]


def test_two_gpu_line_scales():
    d = _load("r2_bench_2gpu.json")
    assert d["n_gpus"] == 2 and d["scaling"] == "weak"
    assert "side stream" in d["details"]["multi_gpu"]
