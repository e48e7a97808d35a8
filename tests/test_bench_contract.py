"""The bench line contract (driver-facing keys) checked on the committed measured lines under profiles/ and on the
pure-Python helpers of bench.py; no GPU needed."""
import importlib.util
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads([ln for ln in f.read().splitlines() if ln.startswith("{")][-1])


def test_default_line_carries_the_contract_keys():
    d = load_line("r2_bench_default.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["dtype"] == "f16c8" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "BASELINE config 2" in d["config"]["workload"] and "model" not in d["config"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and r["frac"] == pytest.approx(r["achieved"] / r["peak"])
    assert r["traffic"] and "not this run" in r["traffic_source"]
    assert r["tensor_pipe_frac"] == pytest.approx(2 * r["frac"])
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and "256 of 10000" in c["sample"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 6e9 and e["d2h_bytes_per_step"] > 0 and 0.9 < e["value"] / d["value"] < 1.0
    assert d["gpu_launches"] > 0 and "sw_power_cap" in d["clocks"]["reasons"]
    p = d["parity"]
    assert p["predictions_per_coalition"] >= 10000 and p["top1_agreement_min"] >= 0.999
    assert p["accuracy_utility_abs_err_max"] <= p["one_sample"] * (1 + 1e-9) and p["shapley_abs_err"] < 1e-3
    assert p["anchor_oracle"]["f16c8"]["top1_agreement"] == 1.0 and p["anchor_f32"]["f16c8"]["top1_agreement_min"] == 1.0
    assert d["throughput_mode"]["precision"] == "f16"


def test_reference_arm_and_multi_gpu_lines():
    r = load_line("r2_bench_reference_arm.json")
    assert r["impl"] == "reference" and r["gpu_launches"] == 0 and r["cpu_baseline"]["kind"] == "reference"
    assert r["e2e"]["h2d_bytes_per_step"] == 0 and r["e2e"]["value"] == r["value"]
    for name, n in (("r2_bench_4gpu.json", 4), ("r2_bench_8gpu.json", 8)):
        d = load_line(name)
        t = d["time_to_shapley"]
        assert d["n_gpus"] == n and t["n_gpus"] == n and t["coalitions"] == 255 and t["scaling"] == "strong"
        assert d["time_to_shapley_s"] == t["value"] <= 1.15 * t["ideal_s"]
        assert t["cross_rank_identity"]["bit_identical"] is True


def test_bench_helpers():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    import argparse

    a = argparse.Namespace(clients=8, vit="base", image=224, val=10000, classes=10, coalition_batch=8, lora_rank=0, lora_path="shared")
    assert "BASELINE config 2" in bench.workload_name(a) and "255 coalitions" in bench.workload_name(a)
    a.lora_rank = 16
    assert "PEFT-LoRA" in bench.workload_name(a)
    peaks, src = bench.load_peaks()
    assert peaks["hbm_gbs"] > 1000 and peaks["bf16_tflops"] > 100 and ("measured" in src or "fallback" in src)
