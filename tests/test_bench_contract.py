"""bench.py contract checks that need no GPU: the JSON lines of the last measured runs (profiles/r2b_bench_products_1gpu.json,
the final library; profiles/r2_bench_products_1gpu.json, first session of round 2) carry every key the driver reads, and the reference arm (`--impl reference`) runs on host cores alone."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _check_common(d):
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["metric"] == "gcn_train_step_ms" and d["unit"] == "ms" and d["higher_is_better"] is False
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32"
    assert "workload" in d["config"] and "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    assert d["cpu_baseline"]["kind"] in ("reference", "port")


import pytest


@pytest.mark.parametrize("name", ["r2b_bench_products_1gpu.json", "r2_bench_products_1gpu.json"])
def test_last_measured_line_has_the_contract_keys(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        d = json.loads(f.read().strip().splitlines()[-1])
    _check_common(d)
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] and 0.5 < r["traffic"] / r["alg_bytes_per_launch"] < 1.5      # ncu DRAM bytes ~ algorithmic bytes
    assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["clocks"]["sm_mhz"] and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["scaling"] == "strong" and d["n_gpus"] == 1 and d["warmup"] >= 3
    # round 2: the CPU leg ran the SAME config (one full step through the port), parity against the exact oracle is green,
    # the e2e leg says it is pipelined, the dense transforms carry their own roofline
    assert d["cpu_baseline"]["same_config"] is True and d["cpu_baseline"]["cores"] >= 1
    assert d["parity"]["ok"] is True and d["parity"]["max_rel_err"] <= 1e-5
    assert d["e2e"]["pipelined"] is True and d["roofline"]["traffic_source"].startswith("static")
    g = d["gemm_roofline"]
    assert 0 < g["frac"] <= 1.0 and g["tf32_peak_tflops_measured"] > 500


def test_reference_arm_runs_on_host_cores():
    """`bench.py --impl reference --config cora`: the REAL reference binary (oracle/_ref/ref_gcn), one dense step"""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_gcn")):
        pytest.skip("oracle/_ref/ref_gcn not built")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cora",
                                   "--steps", "1", "--warmup", "0"], text=True, timeout=600)
    d = json.loads(out.strip().splitlines()[-1])
    _check_common(d)
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["value"] > 0
