#!/usr/bin/env python
"""Per-SASS-instruction stall table of an ncu report (source page): the instructions with the
most warp-stall samples, the stall reason split, and region totals between marker addresses.
    python tools/sass_hot.py rep.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows if len(r) == len(hdr) and r[0] != "Address"]
tot = sum(int(r[ix["Warp Stall Sampling (All Samples)"]] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
agg = {h: sum(int(r[ix[h]] or 0) for r in body) for h in reasons}
print("by reason:", {k[6:]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
print("--- listing (index, samples, executed, top reasons, sass)")
for n, r in enumerate(body):
    s = int(r[ix["Warp Stall Sampling (All Samples)"]] or 0)
    r.append(n)
hot = sorted(body, key=lambda r: -int(r[ix["Warp Stall Sampling (All Samples)"]] or 0))[:top]
for r in sorted(hot, key=lambda r: r[-1]):
    s = int(r[ix["Warp Stall Sampling (All Samples)"]] or 0)
    rs = sorted(((int(r[ix[h]] or 0), h[6:]) for h in reasons), reverse=True)[:3]
    print(f"{r[-1]:5d} {100.0*s/tot:5.2f}% ex={r[ix['Instructions Executed']]:>8s} "
          f"{' '.join(f'{k}:{v}' for v, k in rs if v):40s} {r[1][:90]}")
if len(sys.argv) > 3:
    cuts = [int(c) for c in sys.argv[3].split(",")]
    edges = [0] + cuts + [len(body)]
    for a, b in zip(edges[:-1], edges[1:]):
        seg = body[a:b]
        s = sum(int(r[ix["Warp Stall Sampling (All Samples)"]] or 0) for r in seg)
        ex = sum(int(r[ix["Instructions Executed"]] or 0) for r in seg)
        ag = {h[6:]: sum(int(r[ix[h]] or 0) for r in seg) for h in reasons}
        ag = {k: v for k, v in sorted(ag.items(), key=lambda kv: -kv[1]) if v}
        mma = sum(int(r[ix["Instructions Executed"]] or 0) for r in seg if "DMMA" in r[1])
        f64 = sum(int(r[ix["Instructions Executed"]] or 0) for r in seg
                  if any(t in r[1] for t in ("DADD", "DFMA", "DMUL", "DSETP", "DMNMX")))
        print(f"region [{a},{b}): samples {100.0*s/tot:.1f}% executed {ex} (DMMA {mma}, vector f64 {f64}) {ag}")
