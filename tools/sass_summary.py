#!/usr/bin/env python
"""Instruction counts per kernel from `cuobjdump -sass libflux2b.so` -> profiles/rNN_sass_summary.md.

  python tools/sass_summary.py profiles/r02_sass_summary.md

Counts the mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): UTCHMMA / UTCQMMA / UTCOMMA (tcgen05.mma kind::f16 /
mxf8f6f4 / mxf4nvf4), UTCCP (tcgen05.cp), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor loads / stores), UBLKCP
(bulk copies), UTCBAR (tcgen05.commit), SYNCS (mbarrier), plus HMMA / IMMA (legacy mma.sync: must be absent), MUFU.EX2, and the
total instruction count.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "flux-2-swift-mlx_b200", "libflux2b.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "IMMA",
       "MUFU.EX2", "FFMA2", "STS", "LDS", "BAR"]


def main(dst):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    per = collections.OrderedDict()
    cur = None
    it = iter(names)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next(it, m.group(1))
            cur = re.sub(r"^void ", "", cur)
            cur = re.sub(r"\((int|bool|unsigned int)\)", "", cur)   # template-argument casts: gemm_kernel<(int)256, (bool)1, ...>
            cur = re.sub(r"\(.*$", "", cur).replace("f2b::", "")
            per.setdefault(cur, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        per[cur]["total"] += 1
        for p in PAT:
            if op == p or op.startswith(p + "."):
                per[cur][p] += 1
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    with open(dst, "w") as f:
        f.write(f"# SASS summary of libflux2b.so (cuobjdump -sass, sm_100a), tree at {head}\n\n")
        f.write("Whole library: " + ", ".join(f"{p} {tot[p]}" for p in PAT if tot[p]) + f", instructions {tot['total']}.\n")
        f.write(f"Legacy tensor-core instructions (HMMA / IMMA from mma.sync / wmma): {tot['HMMA'] + tot['IMMA']}.\n\n")
        cols = [p for p in PAT if tot[p]]
        f.write("| kernel | instr | " + " | ".join(cols) + " |\n|---|---:|" + "---:|" * len(cols) + "\n")
        for k, c in sorted(per.items(), key=lambda kv: -(kv[1]["UTCHMMA"] + kv[1]["UTCQMMA"] + kv[1]["UTCOMMA"]) * 100000 - kv[1]["total"]):
            f.write(f"| `{k[:120]}` | {c['total']} | " + " | ".join(str(c[p]) if c[p] else "" for p in cols) + " |\n")
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_summary.md"))
