#!/usr/bin/env python
"""Summarise a F2B_PARITY_LOG (tests/conftest.py rel_l2) into a table: per test the largest measured relative error.
usage: F2B_PARITY_LOG=gpurun_out/parity.log python -m pytest tests -m gpu -q ; python tools/parity_margins.py gpurun_out/parity.log profiles/r02_parity_margins.md"""
import collections
import sys

src, dst = sys.argv[1], sys.argv[2]
worst = collections.OrderedDict()
count = collections.Counter()
for line in open(src):
    name, e = line.rstrip("\n").split("\t")
    e = float(e)
    count[name] += 1
    if name not in worst or e > worst[name]:
        worst[name] = e
groups = collections.OrderedDict()
for name, e in worst.items():
    fn = name.split("[")[0]
    g = groups.setdefault(fn, [0.0, 0, 0])
    g[0] = max(g[0], e); g[1] += 1; g[2] += count[name]
with open(dst, "w") as f:
    f.write("# Measured parity errors of the GPU test-suite (largest rel-L2 per test function over its parameter cases)\n\n")
    f.write("Produced by `F2B_PARITY_LOG=... python -m pytest tests -m gpu` + `tools/parity_margins.py`; every value passed through `rel_l2` in\n"
            "`tests/conftest.py` (device result vs oracle / fp32 reference, or device vs device where 0 is expected). The bound each value is\n"
            "checked against is written in the test itself.\n\n| test | cases | comparisons | largest measured rel-L2 |\n|---|---:|---:|---:|\n")
    for fn, (e, n, c) in groups.items():
        f.write(f"| `{fn}` | {n} | {c} | {e:.3e} |\n")
print(open(dst).read())
