"""CPU: evidence hygiene — every profiles/*.json parses (r01 committed two files with an NCCL banner on line 1), and bench.py's
traffic stamp refuses a profile measured on other kernel sources."""
import glob
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_profile_json_files_parse():
    files = glob.glob(os.path.join(ROOT, "profiles", "*.json"))
    assert files
    for f in files:
        with open(f) as fh:
            json.load(fh)


def test_traffic_is_served_only_for_the_sources_it_was_measured_on(tmp_path, monkeypatch):
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    src = open(os.path.join(ROOT, "bench.py")).read()
    # import without running the fd redirection at module top: evaluate only the helpers
    ns = {"__file__": os.path.join(ROOT, "bench.py")}
    start, end = src.index("ROOT = os.path.dirname"), src.index("class ClockSampler")
    exec("import os, sys, json\n" + src[start:end], ns)
    sha = ns["csrc_sha"]()
    assert len(sha) == 16
    val, why = ns["measured_traffic"]("gemm")
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(path):
        d = json.load(open(path))
        assert (val is not None) == (d.get("csrc_sha") == sha), why
    else:
        assert val is None and "no ncu traffic pass" in why


def test_committed_bench_lines_keep_the_driver_contract():
    """the bench lines under profiles/ (what the driver's own run must look like) carry every key of the bench contract:
    headline, end-to-end leg with its copy sizes, launch count, roofline with measured traffic, CPU baseline, clocks"""
    base = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
            "data", "config", "e2e", "gpu_launches", "roofline", "clocks"}
    for name, n in (("r02_bench.json", 1), ("r02_bench_n2.json", 2), ("r02_bench_n4.json", 4), ("r02_bench_n8.json", 8)):
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        assert base <= set(d), (name, base - set(d))
        assert d["n_gpus"] == n and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
        assert d["metric"] == "dit_steps_per_sec" and d["unit"] == "steps/s" and d["data"] == "synthetic" and d["dtype"] == "bf16"
        assert "workload" in d["config"] and "model" not in d["config"]
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
        assert d["e2e"]["value"] != d["value"] and d["gpu_launches"] > 0
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"]) and d["roofline"]["bound"] == "tensor"
        assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
        assert abs(d["value"] - n * d["steps"] * 4 / (d["ms_per_step"] * d["steps"] * 1e-3)) / d["value"] < 1e-6   # whole-job steps/s
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
        assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"]))
        assert "sp" in d and "configs2" in d and d["sp"]["n_gpus"] == n
        if n == 1:
            assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] == "port"
            assert d["roofline"]["traffic"] is not None   # measured in an ncu pass on these kernel sources
        else:
            for m in d["sp"]["modes"].values():
                assert m["parity_vs_single_gpu"] < 1e-3
    ref = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_ref.json")))
    assert ref["impl"] == "reference" and ref["extrapolated"] is True and ref["e2e"]["h2d_bytes_per_step"] == 0
    assert ref["cpu_baseline"]["kind"] == "port" and ref["metric"] == "dit_steps_per_sec"
