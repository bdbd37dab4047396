"""CPU: the bench lines committed under profiles/ carry every key of the driver's contract, the two arms name the
same workload, and the figures in them hang together (fractions below 1, value = rays / time, e2e below value)."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed")
    with open(path) as f:
        return json.load(f)


@pytest.mark.parametrize("name", ["r2_bench_n1.json", "r2_bench_n2.json", "r2_bench_n4.json", "r2_bench_n8.json"])
def test_own_arm_line(name):
    d = line(name)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "clocks", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["dtype"] == "f32" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.02
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["clocks"]["sm_mhz"] >= 0.9 * d["clocks"]["sm_max_mhz"]
    r = d["rays"]
    traced = r["closest_path"] + r["shadow"] + r["mis_traced"]
    assert abs(traced / (d["ms_per_step"] * d["steps"]) / 1e3 - d["value"]) <= 1e-6 * d["value"]
    if d["n_gpus"] == 1:
        rf = d["roofline"]
        assert rf["bound"] in ("hbm", "tensor") and rf["unit"] == "GB/s" and 0 < rf["frac"] < 1
        assert abs(rf["achieved"] / rf["peak"] - rf["frac"]) < 1e-9
        assert rf["traffic"] is None or rf["traffic"] > 0
        assert 0 < d["roofline_issue"]["frac"] < 1 and 0 < d["roofline_issue"]["lane_frac"] < d["roofline_issue"]["frac"]
        assert 0 < d["l2"]["frac"] < 1
        c = d["cpu_baseline"]
        assert c["kind"] == "reference" and c["cores"] >= 1 and c["value"] > 0 and c["single_thread"]["cores"] == 1
    else:
        assert d["multi_gpu_check"]["ok"] is True and d["multi_gpu_check"]["max_rel_diff"] <= 1e-5
        assert d["collective"]["peer_kernel_ms"] > 0 and d["collective"]["nccl_ms"] > 0
        assert d["other_configs"]["cfg4"]["scaling"] == "strong" and d["other_configs"]["cfg4"]["n_gpus"] == d["n_gpus"]


def test_both_arms_name_the_same_workload():
    own, ref = line("r2_bench_n1.json"), line("r2_bench_reference_arm.json")
    assert ref["impl"] == "reference" and ref["gpu_launches"] == 0
    for key in ("metric", "unit", "higher_is_better", "config"):
        assert own[key] == ref[key], key
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["d2h_bytes_per_step"] == 0 and ref["e2e"]["value"] == ref["value"]
    assert ref["cpu_baseline"]["kind"] == "reference" and ref["cpu_baseline"]["value"] == ref["value"]


def test_weak_scaling_of_the_committed_lines():
    one = line("r2_bench_n1.json")["value"]
    for n in (2, 4, 8):
        eff = line(f"r2_bench_n{n}.json")["value"] / (n * one)
        assert 0.97 <= eff <= 1.03, (n, eff)
