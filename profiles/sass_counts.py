#!/usr/bin/env python
"""Regenerate profiles/r2_sass_counts.md: per kernel of every object of libarreau_b200.so, the count of each tensor-core /
TMEM / TMA mnemonic family in `cuobjdump -sass` (build container, no GPU needed).

    python profiles/sass_counts.py > profiles/r2_sass_counts.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "arreau_b200", "lib")
FAMILIES = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UBLKRED", "UBLKPF", "UTMALDG", "UTMASTG", "UTCCP", "HMMA",
            "SYNCS", "MUFU.TANH", "MUFU.EX2", "CCTL", "LDGSTS", "FFMA2", "HFMA2"]
TITLES = {"model_tc.o": "tcgen05 kernels of the sampling step", "train_net.o": "training forward / backward (tcgen05 kind::tf32 GEMMs)",
          "model_simt.o": "fp32 path, message / fiber / read-out kernels"}

print("# SASS instruction counts per kernel (`cuobjdump -sass arreau_b200/lib/*.o`, sm_100a; `python profiles/sass_counts.py`)\n")
print("For every `.o` of `libarreau_b200.so`, every kernel that contains at least one tensor-core / TMEM / TMA instruction, with "
      "the count of each mnemonic family.\n")
print("`UTCHMMA` = `tcgen05.mma` (kind::f16 and kind::tf32; operand `tmem[..]` = A from tensor memory), `LDTM` / `STTM` = "
      "`tcgen05.ld` / `tcgen05.st`, `UTCBAR` = `tcgen05.commit`, `UTMALDG` = `cp.async.bulk.tensor` (TMA load through a tensor "
      "map), `UBLKCP` = `cp.async.bulk` (TMA engine, no tensor map), `UBLKRED` = `cp.reduce.async.bulk`, `HMMA` = legacy "
      "`mma.sync` (fiber conv 16x16 per channel), `SYNCS` = mbarrier operations.\n")
for obj in sorted(os.listdir(LIB)):
    if not obj.endswith(".o"):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(LIB, obj)], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur.replace("(anonymous namespace)::", "").replace("void ", ""))
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for fam in FAMILIES:
            if re.search(r"\b" + re.escape(fam) + r"\b", line) or (fam + ".") in line:
                kernels[cur][fam] += 1
                if fam == "UTCHMMA" and "tmem[" in line.split("UTCHMMA")[1].split(",")[0]:
                    kernels[cur]["UTCHMMA(A=tmem)"] += 1
    rows = [(k, c) for k, c in kernels.items() if any(c[f] for f in ("UTCHMMA", "UTCQMMA", "LDTM", "UBLKCP", "UTMALDG", "HMMA"))]
    if not rows:
        continue
    cols = FAMILIES + ["UTCHMMA(A=tmem)"]
    print(f"\n## `{obj}` — {TITLES.get(obj, '')}\n")
    print("| kernel | " + " | ".join(cols) + " |")
    print("|---|" + "---:|" * len(cols))
    for k, c in rows:
        print(f"| `{k}` | " + " | ".join(str(c[f]) if c[f] else "" for f in cols) + " |")
