"""The committed bench outputs (profiles/r1_bench_*.json, produced on the B200 box) carry every key
of the bench.py contract; bench.py's argument parser accepts the driver's command lines.  CPU only."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"}


def _load(name):
    p = os.path.join(ROOT, "profiles", name)
    if not os.path.isfile(p):
        pytest.skip("%s not committed yet" % name)
    return json.load(open(p))


def _check_roofline(r):
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0


def test_n1_line_has_the_contract_keys():
    d = _load("r1_bench_n1.json")
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and "workload" in d["config"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != d["value"]                      # measured separately, not a copy
    _check_roofline(d["roofline"])
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    for name, o in d["others"].items():
        assert "error" not in o, (name, o.get("error"))
        _check_roofline(o["roofline"])
    ev = d["others"]["eval-synth"]
    assert ev["roofline"]["bound"] == "tensor" and ev["cpu_baseline"]["bit_identical_to_gpu"] is True


def test_reference_arm_line():
    r = _load("r1_bench_reference_n1.json")
    assert r["impl"] == "reference" and r["value"] > 0
    assert r["e2e"] == {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert r["cpu_baseline"]["value"] == r["value"] and r["cpu_baseline"]["cores"] >= 1
    d = _load("r1_bench_n1.json")
    assert r["metric"] == d["metric"] and r["unit"] == d["unit"] and r["config"]["workload"] == d["config"]["workload"]


def test_n2_lines_are_weak_scaling():
    d1, d2 = _load("r1_bench_n1.json"), _load("r1_bench_n2_eval_synth.json")
    assert d2["n_gpus"] == 2 and d2["scaling"] == "weak"
    one = d1["others"]["eval-synth"]["value"]
    assert 1.6 * one < d2["value"] < 2.4 * one          # users sharded, tables replicated


def test_argument_parser_accepts_the_drivers_command_lines():
    for extra in (["--help"],):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + extra, capture_output=True, text=True)
        assert out.returncode == 0
        for flag in ("--gpus", "--steps", "--warmup", "--impl", "--workload"):
            assert flag in out.stdout + out.stderr
