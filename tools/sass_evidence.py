#!/usr/bin/env python
"""Static SASS counts per kernel of the shipped library (cuobjdump -sass): FP64 tensor pipe
(DMMA), TMA bulk copies (UBLKCP), cp.async (LDGSTS), mbarrier operations (SYNCS).
Usage: python tools/sass_evidence.py > profiles/rN_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cobaya_b200", "lib",
                                                          "libcobaya_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts = collections.defaultdict(collections.Counter)
fn = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    if fn is None:
        continue
    for key in ("DMMA", "UBLKCP", "LDGSTS", "SYNCS", "UTMALDG", "UTCHMMA", "HMMA"):
        if re.search(r"\b" + key, line):
            counts[fn][key] += 1
names = sorted(counts)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout
rows = collections.defaultdict(collections.Counter)
for mangled, d in zip(names, dem.splitlines()):
    short = re.sub(r"\(.*$", "", d).replace("void ", "").replace("(int)", "").replace("(bool)", "")
    rows[short] += counts[mangled]
print("# cuobjdump -sass cobaya_b200/lib/libcobaya_b200.so (sm_100a), instruction counts per kernel")
print("# DMMA.8x8x4 = FP64 tensor pipe (mma.sync.m8n8k4.f64; tcgen05 has no f64 kind); UBLKCP = "
      "cp.async.bulk (TMA bulk copy);")
print("# LDGSTS = cp.async; SYNCS = mbarrier operations.  No UTC*MMA / UTMALDG / HMMA: nothing in "
      "this path is a tcgen05 or tensor-map shape.")
print(f"{'kernel':58s} {'DMMA':>6s} {'UBLKCP':>7s} {'LDGSTS':>7s} {'SYNCS':>6s}")
for k in sorted(rows):
    c = rows[k]
    print(f"{k:58s} {c['DMMA']:6d} {c['UBLKCP']:7d} {c['LDGSTS']:7d} {c['SYNCS']:6d}")
other = sum(c["UTMALDG"] + c["UTCHMMA"] + c["HMMA"] for c in rows.values())
print(f"# UTMALDG + UTCHMMA + HMMA in the whole library: {other}")
