"""The bench line contract (task statement, section 4 + the base contract): checked on the committed line of the headline
run (profiles/r02/bench_cfg2_r2.json -- written by `python bench.py` on a B200, not edited) and on bench.py's static parts.
CPU only."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    p = os.path.join(ROOT, "profiles", "r02", name)
    if not os.path.exists(p):
        pytest.skip(f"{name} not present")
    return json.loads(open(p).read().strip().splitlines()[-1])


@pytest.mark.parametrize("name", ["bench_cfg2_r2.json", "bench_train32k_r2.json", "bench_cfg2_n2.json"])
def test_committed_bench_lines_carry_the_contract_keys(name):
    d = _line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["unit"] == "rays/s" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None, "BASELINE.md publishes no number for this metric"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0.0 < r["frac"] <= 1.0 / 3.0 + 1e-6, "three fp16 products per MAC: 1/3 of the tensor peak is the ceiling"
    # value = whole-job rays per second: consistent with the step time it was derived from
    rays = d["config"].get("rays_total", d["config"].get("rays"))
    assert abs(d["value"] - rays / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    if d["n_gpus"] == 1 and "cpu_baseline" in d:
        c = d["cpu_baseline"]
        for k in ("value", "unit", "cores", "kind", "sample"):
            assert k in c, k
        assert c["kind"] in ("reference", "port")
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert not (bad & set(d["clocks"]["reasons"])), "a throttled run must not be committed as the headline line"


def test_bench_defaults_and_workloads():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert b.METRIC.split(" ")[0] in base["metric"] or "rays/sec" in b.METRIC
    for w in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5", "train32k"):
        assert w in b.WORKLOADS
    cfg2 = b.WORKLOADS["cfg2"]
    assert cfg2["H"] * cfg2["W"] == 1200 * 1600 and cfg2["width"] == 512 and cfg2["n_src"] == 4
    # the sampler degrades to "no samples" without a GPU instead of failing the run
    s = b.ClockSampler(0, period=0.05)
    s.start()
    out = s.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples"}
