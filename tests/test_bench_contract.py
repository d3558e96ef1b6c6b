"""The committed bench outputs (profiles/r2_bench_*.json, produced on the B200 box) carry every key
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
    d = _load("r2_bench_n1.json")
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["scaling"] == "weak"
    assert d["metric"] == "triplets/sec" and "BASELINE configs[4]" in d["config"]["workload"]
    assert d["warmup"] >= 3 and d["gpu_launches"] == d["steps"] and "timed_region" not in d     # one launch per step
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                       # pays the host copies: measured separately, not a copy
    _check_roofline(d["roofline"])
    assert d["roofline"]["kernel"] in ("mf_bpr_sgd_stream_kernel", "mf_bpr_sgd_pipe_kernel") and 0.5 < d["roofline"]["frac"] < 1.1
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert set(d["others"]) >= {"bprmf-ml100k", "neumf-ml100k", "lightgcn-gowalla", "eval-synth"}
    for name, o in d["others"].items():
        assert "error" not in o, (name, o.get("error"))
        _check_roofline(o["roofline"])
        assert o["e2e"]["value"] > 0
    assert "gowalla (29 858 users" in d["others"]["lightgcn-gowalla"]["config"]["workload"]      # the real split
    ev = d["others"]["eval-synth"]
    assert ev["roofline"]["bound"] == "tensor" and ev["cpu_baseline"]["bit_identical_to_gpu"] is True
    assert ev["config"]["ndcg_at_20_all_ranks"] > 0.05                                            # hits are planted


def test_reference_arm_line():
    r = _load("r2_bench_reference_n1.json")
    assert r["impl"] == "reference" and r["value"] > 0
    assert r["e2e"] == {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert r["cpu_baseline"]["value"] == r["value"] and r["cpu_baseline"]["cores"] >= 1
    d = _load("r2_bench_n1.json")
    assert r["metric"] == d["metric"] and r["unit"] == d["unit"] and r["config"]["workload"] == d["config"]["workload"]


def test_multi_gpu_line_communicates():
    d = None
    for n in (8, 4, 2):
        p = os.path.join(ROOT, "profiles", "r2_bench_n%d.json" % n)
        if os.path.isfile(p):
            d = json.load(open(p))
            break
    if d is None:
        pytest.skip("no multi-GPU bench line committed")
    assert d["n_gpus"] == n and d["scaling"] == "weak" and "row-sharded over %d GPU" % n in d["config"]["workload"]
    nv = d["roofline"]["nvlink"]
    assert abs(nv["remote_item_row_fraction"] - (n - 1) / n) < 1e-9 and nv["GBps_per_gpu_per_direction"] > 0
    assert d["config"]["loss_finite"] is True
    ev = d.get("others", {}).get("eval-synth")          # BASELINE: "eval users/sec @1/2/4/8" rides in the same line
    if ev is not None:
        assert ev["unit"] == "users/s" and ev["value"] > 0 and ev["roofline"]["bound"] == "tensor"


def test_argument_parser_accepts_the_drivers_command_lines():
    for extra in (["--help"],):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + extra, capture_output=True, text=True)
        assert out.returncode == 0
        for flag in ("--gpus", "--steps", "--warmup", "--impl", "--workload"):
            assert flag in out.stdout + out.stderr


def test_documents_cite_profiles_that_exist():
    """Every profiles/<file> (or bare r2_*.{json,txt} name) DESIGN.md and profiles/README.md cite is committed."""
    import re
    have = set(os.listdir(os.path.join(ROOT, "profiles")))
    missing = []
    for doc in ("DESIGN.md", os.path.join("profiles", "README.md"), "INTEGRATION.md"):
        text = open(os.path.join(ROOT, doc)).read()
        names = set(re.findall(r"profiles/([A-Za-z0-9_.]+\.(?:json|txt|py|sh|cu))", text))
        names |= set(re.findall(r"`(r[12]_[A-Za-z0-9_.]+\.(?:json|txt|sh))`", text))
        missing += [(doc, n) for n in sorted(names) if n not in have]
    assert not missing, missing
