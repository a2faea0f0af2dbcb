"""CPU-only: the committed bench lines (profiles/r01_bench_*_cfg2.json, written by `python bench.py` on a B200) carry every
key of the measurement contract, and bench.py refuses to run without a GPU (there is no CPU fallback to time)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name):
    return json.load(open(os.path.join(ROOT, "profiles", name)))


def check_common(d):
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "Mpixels*views/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["dtype"] == "f32" and d["vs_baseline"] is None            # BASELINE.md publishes no number for this metric
    assert "workload" in d["config"] and "3111x2074" in d["config"]["workload"]
    assert "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    assert d["roofline"]["bound"] == "hbm" and d["roofline"]["unit"] == "GB/s"
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-3
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    assert d["steps"] >= 1 and d["warmup"] >= 1 and d["value"] > 0 and d["ms_per_step"] > 0


def test_ours_line():
    d = load("r01_bench_ours_cfg2.json")
    check_common(d)
    assert d["impl"] == "ours" and d["warmup"] >= 3
    assert d["gpu_launches"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["e2e"]["value"] < d["value"]                               # host copies inside the timed region
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["roofline"]["traffic"] is None or d["roofline"]["traffic"] > 0


def test_reference_line():
    d = load("r01_bench_reference_cfg2.json")
    check_common(d)
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "reference"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"] == {**load("r01_bench_ours_cfg2.json")["config"]}   # same workload on both arms


def test_bench_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 2
    assert "no CUDA device" in json.loads(r.stdout.strip().splitlines()[-1])["error"]
