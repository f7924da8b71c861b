"""Summarise an ncu source-page CSV: hottest SASS instructions by warp-stall samples, per kernel.
usage: ncu -i rep --page source --csv > src.csv ; python scripts/ncu_hot.py src.csv [topN]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    h = b["hdr"]
    si, ai = h.index("Source"), h.index("Warp Stall Sampling (All Samples)")
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    data = []
    for r in b["rows"]:
        try:
            data.append((int(r[ai]), r))
        except Exception:
            pass
    tot = sum(d[0] for d in data) or 1
    print("===", b["name"][:100], "samples", tot)
    agg = {}
    for n, r in data:
        for i, c in stall_cols:
            try:
                agg[c] = agg.get(c, 0) + int(r[i])
            except Exception:
                pass
    print("   stall mix:", ", ".join(f"{k[6:]} {100 * v / tot:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for n, r in sorted(data, key=lambda d: -d[0])[:top]:
        reasons = sorted(((int(r[i]) if r[i].isdigit() else 0, c[6:]) for i, c in stall_cols), reverse=True)[:2]
        print(f"{n:7d} {100 * n / tot:5.1f}%  {r[si].strip()[:90]:90s} {reasons}")
