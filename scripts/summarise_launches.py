"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals.
usage: python scripts/summarise_launches.py launches.csv [n_steps_in_capture] > profiles/launches_rNN.md
The capture's times are cold-cache and serialised: compare SHARES with bench.py's own table, not absolutes."""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = [r for r in csv.reader(open(path)) if len(r) >= 15 and r[0].isdigit()]
agg = OrderedDict()
tot = 0.0
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").strip()
    name = re.sub(r"<.*", lambda m: m.group(0) if "conv_gemm" in name else "", name)
    ns = float(r[14])
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ns
    tot += ns
print(f"# ncu launch list summary: {path}")
print(f"{len(rows)} launches captured over {steps} step(s) incl. warm-up; total device time {tot / 1e6:.2f} ms\n")
print("| kernel | launches | total ms | share |")
print("|---|---:|---:|---:|")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k[:90]}` | {n} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% |")
