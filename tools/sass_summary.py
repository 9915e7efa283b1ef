#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the tcgen05 / TMEM / TMA path (B200_PROFILING.md):
UTCHMMA (tcgen05.mma kind::f16/tf32), LDTM (tcgen05.ld), UTMALDG / UTMASTG / UTMAREDG (TMA load / store / reduce),
UTCBAR (tcgen05.commit), SYNCS (mbarrier), MULTIMEM, plus HMMA/IMMA (legacy mma.sync, expected 0).

    python tools/sass_summary.py [lib.so] > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "rsl_rl_b200", "lib", "librslrl_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "MULTIMEM", "HMMA", "IMMA", "FFMA", "MUFU"]
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
counts, cur, arch = collections.OrderedDict(), None, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
print(f"# {os.path.relpath(lib, ROOT)}  arch {arch}  ({len(counts)} kernels); columns: " + " ".join(KEYS))
w = max(len(k) for k in counts) if counts else 10
for name, c in counts.items():
    print(f"{name[:110]:110s} " + " ".join(f"{c[k]:7d}" for k in KEYS))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print(f"{'TOTAL':110s} " + " ".join(f"{tot[k]:7d}" for k in KEYS))
