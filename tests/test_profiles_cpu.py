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
