#!/usr/bin/env python
"""Per-kernel totals from an ncu launch list (--metrics gpu__time_duration.sum --csv): count, total us, mean us, share.
usage: python profiles/summarize_launches.py launches.csv [top]"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] in ("ns", "nsecond") else v
    name = re.sub(r"\(.*", "", r[ki])[:90]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {sum(v[0] for v in agg.values())} launches, {tot/1e3:.3f} ms of kernel time")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{t/tot*100:6.2f}%  {t:12.1f} us  n={n:6d}  mean {t/n:9.2f} us  {name}")
