"""The bench lines committed under profiles/ carry every key of the measurement contract (the driver reads the
same keys from a live `python bench.py` run at round end); bench.py itself must parse and expose the flags the
driver passes."""
import ast
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def _line(name):
    return json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])


@pytest.mark.parametrize("name", ["r2_bench_1gpu.json", "r2_bench_2gpu_peer.json", "r2_bench_8gpu_peer.json"])
def test_product_line(name):
    d = _line(name)
    assert BASE_KEYS | {"gpu_launches", "roofline", "cpu_baseline", "clocks"} <= set(d)
    assert d["metric"] == "qmc_diagram_evals_per_sec" and d["unit"] == "diagram_evals/s" and d["dtype"] == "f64"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["warmup"] >= 3
    assert "workload" in d["config"] and "model" not in d["config"]
    # value = units all ranks processed / max-over-ranks device time
    assert d["value"] == pytest.approx(d["diagram_evals_per_step"] / (d["ms_per_step"] * 1e-3), rel=1e-9)
    # N [sum_bare (2n-1)!!] + (n_tau - 2) N 280, order-0 terms evaluated once (exactly) per step
    N = d["n_gpus"] * 1024
    assert d["diagram_evals_per_step"] == 1 + N * (1 + 3 + 15 + 105) + 198 * (1 + N * 280)
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and e["h2d_bytes_per_step"] > 0 < e["d2h_bytes_per_step"]
    assert e["value"] < d["value"]                      # host copies and the host clock are inside the e2e region
    assert d["gpu_launches"] >= d["steps"] * 2          # bare-step kernel + one persistent run kernel per inchworm! run
    assert d["parity"]["pass"] is True and d["parity"]["tolerance"] == 1e-10
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "executed"} <= set(r)
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9) and 0 < r["frac"] < 1
    # the algorithmic fraction is never quoted alone: what the kernel executes and what binds it ride along
    assert 0 < r["executed"]["executed_frac"] < r["frac"] and 0 < r["executed"]["smem_operand_frac"] < 1
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if d["n_gpus"] == 1:
        c = d["cpu_baseline"]
        assert c["kind"] == "port" and c["cores"] >= 1 and c["unit"] == d["unit"] and c["sample"]


def test_reference_arm_line():
    d = _line("r2_bench_reference_arm.json")
    assert BASE_KEYS | {"impl", "cpu_baseline"} <= set(d) and d["impl"] == "reference"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port"
    p = _line("r2_bench_1gpu.json")
    assert (d["metric"], d["unit"], d["higher_is_better"], d["config"]) == (p["metric"], p["unit"], p["higher_is_better"], p["config"])


def test_bench_flags():
    src = open(os.path.join(ROOT, "bench.py")).read()
    ast.parse(src)
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert '"%s"' % flag in src
