#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small, committed text files under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches.md
  python tools/summarize_ncu.py full gpurun_out/prof_full.ncu-rep profiles/r01_top_kernels.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def launches(src, dst):
    lines = [l for l in open(src) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    agg = collections.OrderedDict()
    n = 0
    for row in r:
        v = float(row[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row[ui], 1e-3)
        name = re.sub(r"\(.*", "", row[ki]).replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v; n += 1
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list: {n} launches, {tot / 1e3:.2f} ms of kernel time (cold-cache, serialised: compare shares)\n\n")
        f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1] / 1e3:.3f} | {100 * a[1] / tot:.1f}% | {a[1] / a[0]:.1f} |\n")
    print(open(dst).read())
    # the same durations per bench kernel class (bench.py reports them beside its event-pair timings while the kernel sources match)
    import json, os
    cls = collections.OrderedDict()
    for k, a in agg.items():
        c = ("conv" if re.search(r"gemm_kernel<\d+, \d+, [12]", k) else "gemm" if k.startswith("gemm_kernel") or k.startswith("wq_stage_kernel") else
             "attn" if k.startswith("attn_kernel") else "gemv" if k.startswith("gemv") else "groupnorm" if k.startswith("gn_") else "elem")
        e = cls.setdefault(c, {"launches": 0, "ms": 0.0})
        e["launches"] += a[0]; e["ms"] += a[1] / 1e3
    sha_file = os.path.join(os.path.dirname(os.path.abspath(src)), "csrc_sha.txt")
    out = {"classes": cls, "csrc_sha": open(sha_file).read().strip() if os.path.exists(sha_file) else None,
           "_source": "ncu --metrics gpu__time_duration.sum --clock-control none over every launch of one klein4b 1024x1024 image (cold-cache, serialised)"}
    json.dump(out, open(os.path.splitext(dst)[0] + ".json", "w"), indent=1)


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"# ncu --set full capture: {src}\n")
        for row in rows[2:]:
            f.write(f"\n## `{row[ki][:110]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEEP:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {k} | {row[i]} | {units[i]} |\n")
    print(open(dst).read()[:6000])


def traffic(src, dst):
    """ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over every launch of one image -> mean DRAM bytes per launch
    per kernel class (json read by bench.py for roofline.traffic)."""
    import json
    lines = [l for l in open(src) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, ii, mi, vi, ui = (hdr.index(k) for k in ("Kernel Name", "ID", "Metric Name", "Metric Value", "Metric Unit"))
    per = collections.OrderedDict()
    for row in r:
        if not row[mi].startswith("dram__bytes"):
            continue
        v = float(row[vi].replace(",", ""))
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(row[ui], 1.0)
        name = re.sub(r"\(.*", "", row[ki]).replace("void ", "")
        per.setdefault((row[ii], name), 0.0)
        per[(row[ii], name)] += v
    cls = collections.OrderedDict()
    for (_, name), v in per.items():
        c = ("conv" if re.search(r"gemm_kernel<\d+, \d+, [12]", name) else "gemm" if name.startswith("gemm_kernel") else
             "attn" if name.startswith("attn_kernel") else name)
        a = cls.setdefault(c, [0, 0.0])
        a[0] += 1; a[1] += v
    out = {c: a[1] / a[0] for c, a in cls.items()}
    out["_launches"] = {c: a[0] for c, a in cls.items()}
    import os
    sha_file = os.path.join(os.path.dirname(os.path.abspath(src)), "csrc_sha.txt")
    out["csrc_sha"] = open(sha_file).read().strip() if os.path.exists(sha_file) else None   # kernel sources the pass ran on (bench.py csrc_sha)
    out["_source"] = "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one klein4b 1024x1024 image (bench.py --profile-one); mean bytes per launch"
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1)[:3000])


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
