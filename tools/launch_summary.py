#!/usr/bin/env python
"""Condense an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv) into per-kernel totals: launches, time, share, DRAM bytes.
    python tools/launch_summary.py gpurun_out/launches_r2.csv > profiles/r2_launch_summary.txt"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tu = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
for r in rows:
    k = r["Kernel Name"].split("(")[0].replace("void ", "")
    a = agg.setdefault(k, {"us": 0.0, "rd": 0.0, "wr": 0.0, "ids": set()})
    v = float(r["Metric Value"].replace(",", ""))
    if r["Metric Name"] == "gpu__time_duration.sum":
        a["us"] += v * tu.get(r["Metric Unit"], 1.0)
        a["ids"].add(r["ID"])
    elif r["Metric Name"] == "dram__bytes_read.sum":
        a["rd"] += v * unit[r["Metric Unit"]]
    elif r["Metric Name"] == "dram__bytes_write.sum":
        a["wr"] += v * unit[r["Metric Unit"]]
tot = sum(a["us"] for a in agg.values())
print(f"# {sys.argv[1]}: {len(rows) // 3} launches, {tot / 1e3:.2f} ms of kernel time "
      "(ncu: serialised, cold caches -- shares, not absolutes, are comparable with the bench)")
print(f"{'kernel':58s} {'launches':>8s} {'ms':>9s} {'share':>6s} {'us/launch':>10s} "
      f"{'dram rd MB':>11s} {'dram wr MB':>11s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    n = max(len(a["ids"]), 1)
    print(f"{k[:58]:58s} {n:8d} {a['us'] / 1e3:9.3f} {a['us'] / tot:6.3f} {a['us'] / n:10.1f} "
          f"{a['rd'] / 1e6:11.1f} {a['wr'] / 1e6:11.1f}")
