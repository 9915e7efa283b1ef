#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed here.

  python profiles/summarize_ncu.py launches <launch_list.csv>       # per-kernel share of device time
  python profiles/summarize_ncu.py report <file.ncu-rep> [regex]    # key metrics per captured launch
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] in ("ns", "nsecond") else v
        name = re.sub(r"\(.*", "", r[ki])[-60:]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"total device time {tot:.1f} us over {sum(v[0] for v in agg.values())} launches (ncu: cold caches, serialised)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
        print(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  avg {v[1] / v[0]:8.2f} us  {k}")


def report(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        if pattern and not re.search(pattern, r[ki]):
            continue
        print("==", re.sub(r"\(.*", "", r[ki]))
        for k in KEYS:
            if k in hdr:
                print(f"   {k:85s} {r[hdr.index(k)]} {units[hdr.index(k)]}")


def traffic(path, key, pattern="mlp3_forward|tc_gemm", out="profiles/dense_traffic.json"):
    """DRAM bytes (read + write) of the captured dense launches of ONE minibatch -> profiles/dense_traffic.json[key]
    (read by bench.py for roofline.traffic).  usage: traffic <file.ncu-rep> <config>_<precision> [regex] [out.json]"""
    import json
    import os
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total, n, per = 0.0, 0, []
    for r in rows[2:]:
        if not re.search(pattern, r[ki]):
            continue
        b = float(r[ri].replace(",", "")) * scale[units[ri]] + float(r[wi].replace(",", "")) * scale[units[wi]]
        total += b
        n += 1
        per.append((re.sub(r"\(.*", "", r[ki])[-48:], round(b / 1e6, 2)))
    data = json.load(open(out)) if os.path.exists(out) else {}
    data[key] = dict(bytes_per_minibatch=total, launches_per_minibatch=n, per_launch_mb=per,
                     source=f"ncu --set full, {os.path.basename(path)}")
    json.dump(data, open(out, "w"), indent=1)
    print(key, f"{total / 1e6:.1f} MB over {n} launches", per)


if __name__ == "__main__":
    {"launches": launches, "report": report, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
